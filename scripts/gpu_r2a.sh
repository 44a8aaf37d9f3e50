#!/bin/bash
# round 2, first hardware run: new epilogue (33 passes, seeded boundary terms), L2 hints A/B,
# every stiffness variant at every degree (selection data for launch_stiffness_n)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for h in 0 1; do
  FUS_STAGE_HINTS=$h timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline \
      > gpurun_out/r2a_bench_hints$h.json 2> gpurun_out/r2a_bench_hints$h.err
done
timeout 600 python scripts/bench_sweep.py --degrees 2,3,4,5,6,7 --variants=-1,0,2,3,4,5,6 --geometry-modes 0 \
   --models "" --repeats 20 > gpurun_out/r2a_variants.jsonl 2> gpurun_out/r2a_variants.err
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench_hints0.json | cut -c1-600
cat gpurun_out/r2a_bench_hints1.json | cut -c1-600
