#!/bin/bash
cd "$(dirname "$0")/.."
run() { python tools/bringup_index.py --case $1 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
if 'iters' in r:
    it = r['iters'][-1]
    print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total', 'wall_ms')}, 'flag', it['n_flagged'], 'cg', it['cta_group'])
else:
    print(str(r)[:300])
"; }
echo "== argmax 1M x 64"; run perf_argmax_cg1_1m_q64
echo "== argmax 1M x 64, no epilogue"; AVL_DEBUG_FLAGS=4 run perf_argmax_cg1_1m_q64
echo "== argmax 1M x 64, cg2"; run perf_argmax_cg2_1m_q64
python tools/bringup_index.py --only argmax_ 2>&1 | cut -c1-100
