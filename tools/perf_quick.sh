#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { python tools/bringup_index.py --case $1 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
if 'iters' in r:
    it = r['iters'][-1]
    print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total', 'wall_ms')}, 'cands', it['n_candidates'], 'flag', it['n_flagged'], 'cg', it['cta_group'])
else:
    print(str(r)[:300])
"; }
echo "== topk 4M x 256"; run perf_topk_cg2_4m
echo "== topk 1M x 64"; run perf_topk_cg1_1m_q64
echo "== argmax 1M x 64"; run perf_argmax_cg1_1m_q64
python tools/bringup_build.py 2>&1 | grep perf | cut -c1-400
