#!/bin/bash
# round 2: hardware run of the fused peer transport on N GPUs
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_r2b.sh 2 tag'
N=${1:-2}; TAG=${2:-r2b}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
export FUS_HALO_TIMEOUT_S=10
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt
echo "== multi-GPU parity (NCCL in order, NCCL side stream, fused peer; box + unstructured)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mp_model_check.py > $OUT/${TAG}_mp_check_w$N.log 2>&1
echo "mp check exit $?"; grep "^{" $OUT/${TAG}_mp_check_w$N.log | tail -1 | cut -c1-1700; tail -3 $OUT/${TAG}_mp_check_w$N.log | cut -c1-300
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    par=d.get("parity") or {}
    print(d["n_gpus"], "gpus", "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,3), "e2e", round(d["e2e"]["value"]/1e9,2), "op_ms", round(d["roofline"]["avg_launch_ms"],4), "epi_ms", round(d["roofline"]["stage_epilogue_avg_ms"],4), "parity u", par.get("u_rel_l2"), "apply", par.get("apply_rel_l2"), "halo:", d["config"]["halo"][:24])
except Exception as e:
    print("failed", e)
PY
}
echo "== bench N=1"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; show $OUT/${TAG}_bench_n1.json; tail -2 $OUT/${TAG}_bench_n1.err | cut -c1-300
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n fused peer"
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n${n}.json 2> $OUT/${TAG}_bench_n${n}.err
    show $OUT/${TAG}_bench_n${n}.json; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_n${n}.err | tail -n 3 | cut -c1-300
  fi
done
if [ "${NCCL_TOO:-1}" = "1" ]; then
echo "== bench N=$N NCCL in stream order"
FUS_HALO_TRANSPORT=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-parity > $OUT/${TAG}_bench_n${N}_nccl.json 2> $OUT/${TAG}_bench_n${N}_nccl.err
show $OUT/${TAG}_bench_n${N}_nccl.json
fi
