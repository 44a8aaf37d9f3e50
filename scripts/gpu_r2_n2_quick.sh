#!/bin/bash
# quick 2-GPU check: partitioned parity (three models x three transports) and the N=2 bench line
TAG=${1:-r2n2}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
export FUS_HALO_TIMEOUT_S=10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tests/mp_model_check.py > $OUT/${TAG}_mp_check_w2.log 2>&1
echo "mp check exit $?"; grep "^{" $OUT/${TAG}_mp_check_w2.log | tail -1 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench_n2.json") if l.startswith("{")][0])
print("N=2 ms/step", d["ms_per_step"], "G/s", d["value"]/1e9, "parity", d.get("parity",{}).get("u_rel_l2"), "op", d["roofline"]["avg_launch_ms"], "epi", d["roofline"]["stage_epilogue_avg_ms"], d.get("ms_per_step_by_rank"))
PY
