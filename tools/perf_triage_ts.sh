#!/bin/bash
cd "$(dirname "$0")/.."
for fl in 0 4 5 6 7; do
  echo "== TS AVL_DEBUG_FLAGS=$fl (1=no MMA, 2=no loads, 4=no epilogue)"
  AVL_DEBUG_FLAGS=$fl python tools/bringup_index.py --case perf_topk_ts_4m 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
if 'iters' in r:
    it = r['iters'][-1]
    print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total')}, 'cands', it['n_candidates'])
else:
    print(str(r)[:300])
"
done
