#!/bin/bash
# A/B on one GPU: L2 policy of the G stream (evict-first) with the reversed operator
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(d["config"]["workload"], "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,3), "op_ms", round(d["roofline"]["avg_launch_ms"],4), "epi_ms", round(d["roofline"]["stage_epilogue_avg_ms"],4))
except Exception as e:
    print("failed", e)
PY
}
for ev in 1 0 1 0; do
  for model in linear westervelt; do
    FUS_G_EVICT_FIRST=$ev timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity --model $model > $OUT/r2j_ev${ev}_$model.json 2> $OUT/r2j_ev${ev}_$model.err
    echo -n "g_evict_first=$ev "; show $OUT/r2j_ev${ev}_$model.json
  done
done
for ev in 1 0; do
  FUS_G_EVICT_FIRST=$ev FUS_STAGE_HINTS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity > $OUT/r2j_ev${ev}_hints.json 2> $OUT/r2j_ev${ev}_hints.err
  echo -n "g_evict_first=$ev + epilogue hints "; show $OUT/r2j_ev${ev}_hints.json
done
for ev in 1 0; do
  FUS_G_EVICT_FIRST=$ev timeout 300 python scripts/bench_sweep.py --degrees 4,5,6,7 --variants=-1 --geometry-modes 0 --models "" --repeats 20 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('  sweep ev=$ev P',r['P'],round(r['ms_min'],4),round(r['frac_of_measured_peak'],3))"
done
