#!/bin/bash
# round 2, closing run on ONE GPU with the final library: the GPU suite, the default bench line
# (driver command: extras, parity block, CPU baseline), the reference arm, the profiler evidence.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r2_final1.sh r2z'
TAG=${1:-r2z}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench.py (driver default)"
timeout 900 python bench.py > $OUT/${TAG}_bench_default_1gpu.json 2> $OUT/${TAG}_bench_default_1gpu.err; echo "bench rc=$?"
cut -c1-900 $OUT/${TAG}_bench_default_1gpu.json
echo "== bench.py --impl reference"
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err; echo "reference rc=$?"
cut -c1-600 $OUT/${TAG}_bench_reference_arm.json
echo "== profiler"
bash scripts/gpu_r2_profile.sh ${TAG}p > $OUT/${TAG}_profile.log 2>&1; tail -5 $OUT/${TAG}_profile.log
rm -f $OUT/${TAG}p_prof_stiffness_P4.ncu-rep
