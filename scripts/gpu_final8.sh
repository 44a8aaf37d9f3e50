#!/bin/bash
OUT=gpurun_out; TAG=${1:-r1m}
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(d["n_gpus"], "gpus", d["config"]["workload"], "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,2), "e2e", round(d["e2e"]["value"]/1e9,2), "op_ms", round(d["roofline"]["avg_launch_ms"],4), d["config"]["issue"][:12], "host_enq", round(d["host_enqueue_ms_per_step"],4))
except Exception as e:
    print("failed", e)
PY
}
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 40 --warmup 5 > $OUT/bench_n${n}_${TAG}.json 2> $OUT/bench_n${n}_${TAG}.err
show $OUT/bench_n${n}_${TAG}.json; grep -i "error" $OUT/bench_n${n}_${TAG}.err | head -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --degree 5 --cells 100 --steps 20 --warmup 3 > $OUT/bench_c5_n8_${TAG}.json 2> $OUT/bench_c5_n8_${TAG}.err
show $OUT/bench_c5_n8_${TAG}.json; grep -i "error" $OUT/bench_c5_n8_${TAG}.err | head -3
