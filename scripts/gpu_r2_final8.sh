#!/bin/bash
# round 2, closing run on EIGHT GPUs with the final library: weak scaling N=8, 4 and BASELINE config 5
TAG=${1:-r2t}
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
for n in 8 4; do
  echo "== bench N=$n (driver command without the child runs)"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n${n}.json 2> $OUT/${TAG}_bench_n${n}.err
  show $OUT/${TAG}_bench_n${n}.json; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_n${n}.err | tail -n 3 | cut -c1-300
done
echo "== BASELINE config 5: P=5, 100^3 cells per GPU, 8 GPUs (the child run of the N=8 extras)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --degree 5 --cells 100 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --parity-steps 5 > $OUT/${TAG}_bench_c5_n8.json 2> $OUT/${TAG}_bench_c5_n8.err
show $OUT/${TAG}_bench_c5_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_c5_n8.err | tail -n 3 | cut -c1-300
echo "== N=1 on the same box"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; show $OUT/${TAG}_bench_n1.json
