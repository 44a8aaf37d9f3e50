#!/bin/bash
# round 2: the cell-per-thread kernel for P = 2 (stiffness_variant 7) against the column kernel
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2l.sh r2l'
TAG=${1:-r2l}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "== parity of the new kernel"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cell_per_thread" 2>&1 | tail -5
echo "== P=2 sweep: column kernel, cell kernel"
timeout 300 python scripts/bench_sweep.py --degrees 2 --variants=0,7 --models "" > $OUT/${TAG}_sweep_P2.jsonl 2> $OUT/${TAG}_sweep_P2.err; echo "exit $?"
for cb in 2 11 14; do
  timeout 300 python scripts/bench_sweep.py --degrees 2 --variants=7 --col-blocks-per-sm $cb --models "" >> $OUT/${TAG}_sweep_P2.jsonl 2>> $OUT/${TAG}_sweep_P2.err; echo "cb $cb exit $?"
done
python - <<PY
import json
for l in open("$OUT/${TAG}_sweep_P2.jsonl"):
    d=json.loads(l); print(d["P"], "variant", d["variant"], "blocks", d["col_blocks_per_sm"], "ms", round(d["ms_min"],4), round(d["ms_median"],4), "frac", round(d["frac_of_measured_peak"],3), "err", d["rel_l2_vs_first_config"])
PY
echo "== ncu of the cell kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stiffness_cell -s 2 -c 1 \
    -f -o $OUT/${TAG}_prof_cell_P2 python scripts/bench_sweep.py --degrees 2 --variants=7 --models "" --repeats 3 > $OUT/${TAG}_ncu_cell.log 2>&1
echo "ncu exit $?"
python scripts/ncu_digest.py $OUT/${TAG}_prof_cell_P2.ncu-rep --stalls > $OUT/${TAG}_ncu_full_stiffness_cell_P2.json
ncu -i $OUT/${TAG}_prof_cell_P2.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_source_hot.py 40 > $OUT/${TAG}_ncu_source_hot_cell_P2.txt
ls -la $OUT/${TAG}_prof_cell_P2.ncu-rep
