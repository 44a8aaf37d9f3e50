#!/bin/bash
# round 2, closing run on TWO GPUs: multi-GPU tests and the driver's N=2 command (with its extras)
TAG=${1:-r2u}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
export FUS_HALO_TIMEOUT_S=10
echo "== pytest tests/test_gpu_multi.py"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest_multi.log
echo "== driver command, N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_n2_full.json 2> $OUT/${TAG}_bench_n2_full.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench_n2_full.json") if l.startswith("{")][0])
print("N=2 ms/step", d["ms_per_step"], "G/s", d["value"]/1e9, "parity", d.get("parity",{}).get("u_rel_l2"), "op", d["roofline"]["avg_launch_ms"], "epi", d["roofline"]["stage_epilogue_avg_ms"])
ex=d.get("extras",{})
print("extras keys", list(ex.keys()))
print(json.dumps(ex.get("baseline_configs",{}))[:1500])
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/${TAG}_bench_n2_full.err | tail -5 | cut -c1-300
