#!/bin/bash
cd "$(dirname "$0")/../.."
for slab in "" "103,128" "0,54"; do
  echo "== prepared, 240 frames, slab='$slab' batch=8: $(AVL_FRAMES=240 AVL_PREPARED=1 AVL_SLAB=$slab AVL_BATCH=8 python tools/perf_build.py 2>/dev/null | tail -1)"
done
