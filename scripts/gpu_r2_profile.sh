#!/bin/bash
# round 2: profiler evidence of the final build (1 GPU)
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_profile.sh r2p'
TAG=${1:-r2p}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --no-extras --no-parity --no-cpu-baseline"
echo "== the profiled command on its own"
timeout 300 $BENCH > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err; echo "exit $?"
echo "== launch list (gpu__time_duration per launch)"
FUS_USE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_ncu_launches_bench_P4.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1; echo "exit $?"
echo "== ncu --set full: stiffness kernel and epilogues of the headline step"
FUS_USE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_line -s 6 -c 1 \
    -f -o $OUT/${TAG}_prof_stiffness_P4 $BENCH > $OUT/${TAG}_ncu_stiffness.log 2>&1; echo "exit $?"
python scripts/ncu_digest.py $OUT/${TAG}_prof_stiffness_P4.ncu-rep --stalls > $OUT/${TAG}_ncu_full_stiffness_P4.json
FUS_USE_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stage -s 8 -c 4 \
    -f -o $OUT/${TAG}_prof_stage_P4 $BENCH > $OUT/${TAG}_ncu_stage.log 2>&1; echo "exit $?"
python scripts/ncu_digest.py $OUT/${TAG}_prof_stage_P4.ncu-rep > $OUT/${TAG}_ncu_full_rk4_stage_P4.json
rm -f $OUT/${TAG}_prof_stage_P4.ncu-rep     # gpurun brings back at most 64 MiB: digests travel, one report
echo "== ncu --set full: the kernels of the degree sweep that sit below 0.85 of the HBM peak"
for P in 2 6 7; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_ -s 2 -c 1 \
      -f -o $OUT/${TAG}_prof_stiffness_P${P} python scripts/bench_sweep.py --degrees $P --variants=-1 \
      --geometry-modes 0 --models "" --repeats 3 > $OUT/${TAG}_ncu_P${P}.log 2>&1
  echo "ncu P=$P exit $?"
  python scripts/ncu_digest.py $OUT/${TAG}_prof_stiffness_P${P}.ncu-rep --stalls > $OUT/${TAG}_ncu_full_stiffness_P${P}.json
  ncu -i $OUT/${TAG}_prof_stiffness_P${P}.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_source_hot.py 25 > $OUT/${TAG}_ncu_source_hot_P${P}.txt
  rm -f $OUT/${TAG}_prof_stiffness_P${P}.ncu-rep
done
ncu -i $OUT/${TAG}_prof_stiffness_P4.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_source_hot.py 25 > $OUT/${TAG}_ncu_source_hot_P4.txt
ls -la $OUT | grep ${TAG}
