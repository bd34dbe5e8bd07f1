#!/bin/bash
cd "$(dirname "$0")/.."
run() { python tools/bringup_index.py --case $1 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
if 'iters' in r:
    it = r['iters'][-1]
    print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total')}, 'cands', it['n_candidates'], 'cg', it['cta_group'])
else:
    print(str(r)[:300])
"; }
echo "== TS"; run perf_topk_ts_4m
echo "== TS flags 7"; AVL_DEBUG_FLAGS=7 run perf_topk_ts_4m
echo "== TS flags 6"; AVL_DEBUG_FLAGS=6 run perf_topk_ts_4m
echo "== SS"; AVL_TS=0 run perf_topk_ts_4m
echo "== SS flags 7"; AVL_TS=0 AVL_DEBUG_FLAGS=7 run perf_topk_ts_4m
echo "== argmax C2"; run perf_argmax_cg1_1m_q64
python tools/bringup_index.py --only topk_ts,topk_cg2,argmax_cg2_q256,screen_cg2 2>&1 | cut -c1-120
