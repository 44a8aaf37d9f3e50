#!/bin/bash
# Multi-GPU call of the next round (8 GPUs): weak scaling of the headline workload with the default
# block decomposition and with slabs (each rank then has at most two neighbours, faces only), the
# multi-GPU parity check, config 5.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_next_round_multi.sh r2m'
OUT=gpurun_out; TAG=${1:-r2m}
mkdir -p $OUT
run() { # name nproc extra-args...
  local name=$1 n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 40 --warmup 5 "$@" \
      > $OUT/bench_${name}_${TAG}.json 2> $OUT/bench_${name}_${TAG}.err
  python - "$OUT/bench_${name}_${TAG}.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(d["n_gpus"], d["config"]["process_grid"], "ms/step", round(d["ms_per_step"], 4), "G/s",
          round(d["value"] / 1e9, 2))
except Exception as e:
    print("failed", sys.argv[1], e)
PY
}
echo "== multi-GPU parity check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
    --master-port 29555 tests/mp_model_check.py > $OUT/mp_check_${TAG}.log 2>&1; echo "mp check exit $?"
tail -n 5 $OUT/mp_check_${TAG}.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_n1_${TAG}.json 2> $OUT/bench_n1_${TAG}.err
for n in 2 4 8; do run n${n}_blocks $n; done
run n4_slabs 4 --pgrid 4,1,1
run n8_slabs 8 --pgrid 8,1,1
run n8_slabs_yz 8 --pgrid 1,2,4
run c5_n8 8 --degree 5 --cells 100
# BASELINE config 4: Westervelt, P=4, 1/2/4/8 GPUs
python bench.py --model westervelt --steps 40 --warmup 5 --no-cpu-baseline --no-extras \
    > $OUT/bench_westervelt_n1_${TAG}.json 2> $OUT/bench_westervelt_n1_${TAG}.err
for n in 2 4 8; do run westervelt_n${n} $n --model westervelt; done
ls -la $OUT | tail -n 12
