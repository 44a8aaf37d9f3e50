#!/bin/bash
# quick check: degree sweep with the library's kernel choice + a short headline bench
TAG=${1:-r2x}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/bench_sweep.py --degrees 4,5,6,7 --variants=-1 --models "" > $OUT/${TAG}_sweep_auto.jsonl 2> $OUT/${TAG}_sweep_auto.err; echo "exit $?"
python - <<PY
import json
for l in open("$OUT/${TAG}_sweep_auto.jsonl"):
    d=json.loads(l); print(d["P"], "variant", d["variant"], "ms", round(d["ms_min"],4), round(d["ms_median"],4), "frac", round(d["frac_of_measured_peak"],3))
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-parity --no-cpu-baseline > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/${TAG}_bench_short.json") if l.startswith("{")][0])
print("ms/step", d["ms_per_step"], "op", d["roofline"]["avg_launch_ms"], "epi", d["roofline"]["stage_epilogue_avg_ms"])
PY
