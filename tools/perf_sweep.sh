#!/bin/bash
# prefetch-distance sweep of the headline kernel + correctness of everything
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for pf in 0 1 2 4 8; do
  echo "== AVL_PREFETCH_TILES=$pf"
  AVL_PREFETCH_TILES=$pf python tools/bringup_index.py --case perf_topk_cg2_4m 2>&1 | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
it = r['iters'][-1]
print({k: round(it[k], 4) for k in ('wall_ms', 'ms_screen', 'ms_total')}, 'TF', round(r['tflops'], 1), 'GB/s', round(r['gbs'], 1))
"
done
