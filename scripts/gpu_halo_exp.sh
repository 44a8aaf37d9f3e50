#!/bin/bash
# halo experiments on N GPUs: bash scripts/gpu_halo_exp.sh N
N=${1:-4}
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/halo_exp_$tag.err | grep "^{" > gpurun_out/halo_exp_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/halo_exp_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],4), "host_enq", round(d["host_enqueue_ms_per_step"],4), "stiff", round(d["roofline"]["avg_launch_ms"],4), "epi", round(d["roofline"]["stage_epilogue_avg_ms"],4), "G/s", round(d["value"]/1e9,2))
except Exception as e:
    print("$tag failed", e)
PY
}
run default FUS_DUMMY=1
run no_overlap FUS_HALO_OVERLAP=0
run reserve16 FUS_HALO_RESERVE=16
run p2pch2 NCCL_MAX_P2P_NCHANNELS=2 NCCL_MIN_P2P_NCHANNELS=1
run p2pch2_res8 NCCL_MAX_P2P_NCHANNELS=2 NCCL_MIN_P2P_NCHANNELS=1 FUS_HALO_RESERVE=8
