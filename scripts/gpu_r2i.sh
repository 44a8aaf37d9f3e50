#!/bin/bash
# A/B on one GPU: operator walking the cells backwards inside the RK4 loop (L2 reuse with the epilogue)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    par=d.get("parity") or {}
    print(d["config"]["workload"], "ms/step", round(d["ms_per_step"],4), "G/s", round(d["value"]/1e9,3), "op_ms", round(d["roofline"]["avg_launch_ms"],4), "epi_ms", round(d["roofline"]["stage_epilogue_avg_ms"],4), "parity u", par.get("u_rel_l2"))
except Exception as e:
    print("failed", e)
PY
}
for rev in 1 0 1 0; do
  for model in linear westervelt; do
    FUS_REVERSE_OPERATOR=$rev timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity --model $model > $OUT/r2i_rev${rev}_$model.json 2> $OUT/r2i_rev${rev}_$model.err
    echo -n "reverse=$rev "; show $OUT/r2i_rev${rev}_$model.json
  done
done
for P in 5 6; do for rev in 1 0; do
  n=$([ $P = 5 ] && echo 43 || echo 36)
  FUS_REVERSE_OPERATOR=$rev timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity --degree $P --cells $n > $OUT/r2i_rev${rev}_P$P.json 2> $OUT/r2i_rev${rev}_P$P.err
  echo -n "P=$P reverse=$rev "; show $OUT/r2i_rev${rev}_P$P.json
done; done
