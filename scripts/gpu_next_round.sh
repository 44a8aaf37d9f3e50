#!/bin/bash
# First GPU call of the next round (1 GPU): everything this round's last hours could not run.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_next_round.sh r2a'
# 1. full GPU suite (incl. the tests written after the budget ran out: reference 2-D mesh, C ABI from C,
#    float operator classes), smoke, bench (with the child-process sweep: degree sweep x geometry modes)
# 2. the line kernel's pipeline variants on the three models + ncu of variant 5
# 2b. ncu --set full of the on-the-fly geometry kernel (mode 2) to confirm the latency-bound reading
# 3. ncu --set full of the streamed kernel at P=6 and P=7
# 4. BASELINE config 5 (1.0 G dofs, P=5) on ONE GPU with a lean context
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu_${TAG}.log
tail -n 15 $OUT/pytest_gpu_${TAG}.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 | tee $OUT/smoke_${TAG}.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
echo "bench exit $?"; tail -c 2500 $OUT/bench_${TAG}.json; tail -n 3 $OUT/bench_${TAG}.err
echo "== probe (all degrees, modes 0 and 2)"
timeout 300 python scripts/probe_geometry_modes.py 2:107 3:71 4:54 5:43 6:36 7:31 > $OUT/probe_${TAG}.jsonl 2> $OUT/probe_${TAG}.err
cat $OUT/probe_${TAG}.jsonl
echo "== line-kernel pipeline variants (stiffness_variant 3..6 vs the default 2): headline, lossy, Westervelt"
for MODEL in linear lossy westervelt; do
  for V in 2 3 4 5 6; do
    FUS_STIFFNESS_VARIANT=$V timeout 300 python bench.py --model $MODEL --steps 20 --warmup 3 --no-cpu-baseline \
        --no-extras > $OUT/bench_${MODEL}_variant${V}_${TAG}.json 2> $OUT/bench_${MODEL}_variant${V}_${TAG}.err
    python - "$OUT/bench_${MODEL}_variant${V}_${TAG}.json" $MODEL $V <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(sys.argv[2], "variant", sys.argv[3], "ms/step", round(d["ms_per_step"], 4), "operator ms",
          round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("failed", sys.argv[1], e)
PY
  done
done
echo "== ncu full, variants 5 and 6 at P=4 and P=6 (is the loop-end stall gone?)"
for P in 4 6; do
  FUS_STIFFNESS_VARIANT=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_line \
      -s 2 -c 1 -f -o $OUT/prof_stiffness_variant6_P${P}_${TAG} python scripts/bench_sweep.py --degrees $P \
      --variants=6 --geometry-modes 0 --models "" --repeats 3 > $OUT/ncu_variant6_P${P}_${TAG}.log 2>&1
  echo "ncu variant 6 P=$P exit $?"
  FUS_STIFFNESS_VARIANT=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_line \
      -s 2 -c 1 -f -o $OUT/prof_stiffness_variant5_P${P}_${TAG} python scripts/bench_sweep.py --degrees $P \
      --variants=5 --geometry-modes 0 --models "" --repeats 3 > $OUT/ncu_variant5_P${P}_${TAG}.log 2>&1
  echo "ncu variant 5 P=$P exit $?"
done
echo "== ncu full, mode-2 kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stiffness_line -s 6 -c 2 \
    -f -o $OUT/prof_stiffness_mode2_${TAG} python bench.py --steps 2 --warmup 1 --geometry-mode 2 \
    --no-cpu-baseline --no-extras > $OUT/ncu_mode2_${TAG}.log 2>&1
echo "ncu exit $?"
echo "== ncu full, streamed kernels at P=6 and P=7 (0.70 of the HBM peak in round 1: find the stall)"
for P in 6 7; do
  timeout 300 python scripts/bench_sweep.py --degrees $P --variants=-1 --geometry-modes 0 --models "" --repeats 3 \
      > $OUT/sweep_P${P}_${TAG}.jsonl 2> $OUT/sweep_P${P}_${TAG}.err &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_line -s 2 -c 1 \
      -f -o $OUT/prof_stiffness_P${P}_${TAG} python scripts/bench_sweep.py --degrees $P --variants=-1 \
      --geometry-modes 0 --models "" --repeats 3 > $OUT/ncu_P${P}_${TAG}.log 2>&1
  echo "ncu P=$P exit $?"
done
echo "== config 5 (1.0 G dofs, P=5, 200^3 cells) on one GPU, lean context"
# host side of a 1.0 G-dof mesh: ~7 GB dofmap, 4 x 8 GB boundary vectors, pinned state copies: ~100 GB
AVAIL_GB=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
LIMIT=$(cat /sys/fs/cgroup/memory.max 2>/dev/null || echo max)
echo "host memory available: ${AVAIL_GB} GB, cgroup limit: ${LIMIT}"
if [ "$AVAIL_GB" -ge 250 ] && { [ "$LIMIT" = "max" ] || [ "$LIMIT" -ge 268435456000 ]; }; then
  timeout 1200 python bench.py --degree 5 --cells 200 --lean --steps 5 --warmup 2 --no-cpu-baseline --no-extras \
      > $OUT/bench_c5_lean_1gpu_${TAG}.json 2> $OUT/bench_c5_lean_1gpu_${TAG}.err
  echo "config-5 lean exit $?"; tail -c 1500 $OUT/bench_c5_lean_1gpu_${TAG}.json; tail -n 3 $OUT/bench_c5_lean_1gpu_${TAG}.err
elif [ "$AVAIL_GB" -lt 130 ]; then
  echo "skipped: not enough host memory for a 0.5-1.0 G-dof mesh on the host side"
else
  echo "not enough host memory for the 1.0 G-dof host arrays; running 160^3 cells (0.51 G dofs) instead"
  timeout 1200 python bench.py --degree 5 --cells 160 --lean --steps 5 --warmup 2 --no-cpu-baseline --no-extras \
      > $OUT/bench_c5_lean_1gpu_${TAG}.json 2> $OUT/bench_c5_lean_1gpu_${TAG}.err
  echo "160^3 lean exit $?"; tail -c 1500 $OUT/bench_c5_lean_1gpu_${TAG}.json; tail -n 3 $OUT/bench_c5_lean_1gpu_${TAG}.err
fi
ls -la $OUT | tail -n 12
