#!/bin/bash
# round 2: 8-GPU check after the epilogue's dynamic chunk scheduling / parallel flags
TAG=${1:-r2f}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
export FUS_HALO_TIMEOUT_S=10
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    par=d.get("parity") or {}
    print(d["n_gpus"], "gpus", d["config"]["workload"], "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,3), "e2e", round(d["e2e"]["value"]/1e9,2), "op_ms", round(d["roofline"]["avg_launch_ms"],4), "epi_ms", round(d["roofline"]["stage_epilogue_avg_ms"],4), "parity u", par.get("u_rel_l2"), "by rank", [round(x,3) for x in d.get("ms_per_step_by_rank",[])])
except Exception as e:
    print("failed", e)
PY
}
echo "== multi-GPU parity on 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/mp_model_check.py > $OUT/${TAG}_mp_check_w8.log 2>&1
echo "mp check exit $?"; grep "^{" $OUT/${TAG}_mp_check_w8.log | tail -1 | cut -c1-200
for n in 8 4 2; do
  echo "== bench N=$n fused peer"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n${n}.json 2> $OUT/${TAG}_bench_n${n}.err
  show $OUT/${TAG}_bench_n${n}.json; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_n${n}.err | tail -n 3 | cut -c1-300
done
echo "== bench N=1"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; show $OUT/${TAG}_bench_n1.json
