#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_topk.csv python tools/bringup_index.py --case perf_topk_cg2_4m > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_topk.csv') if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    u=row['Metric Unit']; v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    seq.append((row['Kernel Name'].split('(')[0][-40:], v))
idx=[i for i,(n,_) in enumerate(seq) if 'topk_finalize' in n]
i=idx[-1]
for n,v in seq[i-6:i+1]: print(f"{v:10.1f} us  {n}")
PY
