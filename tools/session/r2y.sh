#!/bin/bash
cd "$(dirname "$0")/../.."
for slab in "" "0,128" "103,128" "0,54"; do
  for batch in 1 8; do
    echo "== slab='$slab' batch=$batch: $(AVL_SLAB=$slab AVL_BATCH=$batch python tools/perf_build.py 2>/dev/null | tail -1)"
  done
done
AVL_SLAB="103,128" AVL_BATCH=8 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2y_launches_slab.csv python tools/perf_build.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2y_launches_slab.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows[-60:]: agg[r[4][:44]].append(float(r[-1]))
for k,v in agg.items(): print(k.ljust(46), len(v), round(sum(v)/len(v)/1000,2),'us per launch (8 frames)')
PY
