#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== gpu index tests, AVL_PDL=0"; AVL_PDL=0 python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2j_ab.json --variants "pdl:" "nopdl:AVL_PDL=0"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-build --no-cpu --no-extra 2>/dev/null | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2j_launches_step.csv python tools/perf_screen.py --child 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_step.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-13:]: print(r[4][:50].ljust(50), r[8], r[-1])
PY
