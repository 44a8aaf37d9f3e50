#!/bin/bash
# One gpurun call: GPU parity tests, bench, then (only if the plain bench exited 0) the ncu passes.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_${TAG}.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu_${TAG}.log
tail -n 25 $OUT/pytest_gpu_${TAG}.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5 | tee $OUT/smoke_${TAG}.log
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
BRC=$?
echo "bench exit $BRC"; tail -c 3000 $OUT/bench_${TAG}.json; tail -n 5 $OUT/bench_${TAG}.err
if [ $BRC -eq 0 ] && [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $OUT/plain_${TAG}.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $OUT/ncu_list_${TAG}.log 2>&1
  echo "ncu list exit $?"
  echo "== ncu full (stiffness kernel)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:stiffness_line -s 4 -c 2 \
      -f -o $OUT/prof_stiffness_${TAG} python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $OUT/ncu_full_${TAG}.log 2>&1
  echo "ncu full exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk4_stage -s 4 -c 2 \
      -f -o $OUT/prof_stage_${TAG} python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $OUT/ncu_full2_${TAG}.log 2>&1
  echo "ncu stage exit $?"
fi
ls -la $OUT | tail -n 20
