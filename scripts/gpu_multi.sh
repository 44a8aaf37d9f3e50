#!/bin/bash
# Multi-GPU call: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_multi.sh N tag'
N=${1:-2}; TAG=${2:-r1}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus_${TAG}.txt; free -g | head -2 | tail -1; nproc
echo "== multi-GPU parity"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mp_model_check.py > $OUT/mp_check_${TAG}.log 2>&1
echo "mp check exit $?"; grep "^{" $OUT/mp_check_${TAG}.log | cut -c1-300
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(d["n_gpus"], "gpus", d["config"]["workload"], "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,2), "e2e", round(d["e2e"]["value"]/1e9,2), "op_ms", round(d["roofline"]["avg_launch_ms"],4), "halo:", d["config"]["halo"][:20])
except Exception as e:
    print("failed", e)
PY
}
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > $OUT/bench_n1_${TAG}.json 2> $OUT/bench_n1_${TAG}.err; show $OUT/bench_n1_${TAG}.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 40 --warmup 5 > $OUT/bench_n${n}_${TAG}.json 2> $OUT/bench_n${n}_${TAG}.err
    show $OUT/bench_n${n}_${TAG}.json; tail -n 2 $OUT/bench_n${n}_${TAG}.err | cut -c1-200
  fi
done
if [ "${CONFIG5:-0}" = "1" ]; then
  echo "== config 5: P=5, 100^3 cells per GPU"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --degree 5 --cells 100 --steps 20 --warmup 3 > $OUT/bench_c5_n${N}_${TAG}.json 2> $OUT/bench_c5_n${N}_${TAG}.err
  show $OUT/bench_c5_n${N}_${TAG}.json; tail -n 2 $OUT/bench_c5_n${N}_${TAG}.err | cut -c1-200
fi
