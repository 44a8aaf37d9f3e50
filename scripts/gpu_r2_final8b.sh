#!/bin/bash
# round 2: the driver's own N=8 command, child runs included (Westervelt on 8 GPUs, config 5 on 8 and on 1)
TAG=${1:-r2s8}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
export FUS_HALO_TIMEOUT_S=10
T0=$(date +%s)
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/${TAG}_bench_n8_full.json 2> $OUT/${TAG}_bench_n8_full.err; echo "rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench_n8_full.json") if l.startswith("{")][0])
print("N=8 ms/step", d["ms_per_step"], "G/s", d["value"]/1e9, "parity", d.get("parity",{}).get("u_rel_l2"))
ex=d.get("extras",{}).get("baseline_configs",{})
for k,v in ex.items():
    if isinstance(v,dict): print(k, {kk:v.get(kk) for kk in ("value","ms_per_step","n_gpus","wall_s","exit","skipped","stderr_tail")})
    else: print(k, v)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_n8_full.err | tail -5 | cut -c1-300
