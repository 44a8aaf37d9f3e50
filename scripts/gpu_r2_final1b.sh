#!/bin/bash
# round 2, closing run on ONE GPU after the spill fix: degree sweep with the library's choice, the
# stiffness-variant tests, default bench line, reference arm, profiler evidence (stamps traffic.json)
TAG=${1:-r2y}
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "== degree sweep (library's choice) "
timeout 300 python scripts/bench_sweep.py --degrees 2,3,4,5,6,7 --variants=-1 --models "" > $OUT/${TAG}_sweep_auto.jsonl 2> $OUT/${TAG}_sweep_auto.err; echo "exit $?"
python - <<PY
import json
for l in open("$OUT/${TAG}_sweep_auto.jsonl"):
    d=json.loads(l); print(d["P"], "variant", d["variant"], "ms", round(d["ms_min"],4), round(d["ms_median"],4), "frac", round(d["frac_of_measured_peak"],3))
PY
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== bench.py (driver default)"
timeout 900 python bench.py > $OUT/${TAG}_bench_default_1gpu.json 2> $OUT/${TAG}_bench_default_1gpu.err; echo "bench rc=$?"
cut -c1-400 $OUT/${TAG}_bench_default_1gpu.json
echo "== profiler"
bash scripts/gpu_r2_profile.sh ${TAG}p > $OUT/${TAG}_profile.log 2>&1; tail -3 $OUT/${TAG}_profile.log
rm -f $OUT/${TAG}p_prof_stiffness_P4.ncu-rep
