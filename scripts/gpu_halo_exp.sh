#!/bin/bash
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 40 --warmup 5 2>gpurun_out/halo_exp_$tag.err | grep "^{" > gpurun_out/halo_exp_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/halo_exp_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],4), "host_enq", round(d["host_enqueue_ms_per_step"],4), "stiff", round(d["roofline"]["avg_launch_ms"],4), "epi", round(d["roofline"]["stage_epilogue_avg_ms"],4), "G/s", round(d["value"]/1e9,2))
except Exception as e:
    print("$tag failed", e)
PY
}
run nccl_seq FUS_HALO_TRANSPORT=nccl
run peer_r4 FUS_HALO_TRANSPORT=peer
run peer_r0 FUS_HALO_TRANSPORT=peer FUS_HALO_RESERVE=0
run peer_r8 FUS_HALO_TRANSPORT=peer FUS_HALO_RESERVE=8
run peer_r4b FUS_HALO_TRANSPORT=peer
run nccl_seq_b FUS_HALO_TRANSPORT=nccl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mp_model_check.py 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['status'], {k:v for k,v in d.items() if 'peer' in k})"
