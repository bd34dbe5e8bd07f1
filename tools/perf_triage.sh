#!/bin/bash
# component triage of the headline kernel: which of TMA stream / MMA / epilogue paces it
cd "$(dirname "$0")/.."
for fl in 0 4 1 2 5 6 3 7; do
  echo "== AVL_DEBUG_FLAGS=$fl (1=no MMA, 2=no A loads, 4=no epilogue)"
  AVL_DEBUG_FLAGS=$fl python tools/bringup_index.py --case perf_topk_cg2_4m 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
if 'iters' in r:
    it = r['iters'][-1]
    print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total')}, 'cands', it['n_candidates'])
else:
    print(str(r)[:300])
"
done
echo "== argmax C2 after rerank change"
python tools/bringup_index.py --case perf_argmax_cg1_1m_q64 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1]); it = r['iters'][-1]
print({k: round(it[k], 4) for k in ('ms_screen', 'ms_total', 'wall_ms')}, it['n_flagged'])"
python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -3
