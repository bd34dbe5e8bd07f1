#!/bin/bash
cd "$(dirname "$0")/../.."
echo "== build tests"; python -m pytest tests/test_build_gpu.py tests/test_handoff_p2p_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -2
for b in 8 16; do echo "== prepared, 240 frames, full grid, batch=$b: $(AVL_FRAMES=240 AVL_PREPARED=1 AVL_BATCH=$b python tools/perf_build.py 2>/dev/null | tail -1)"; done
echo "== prepared, slab 103,128 batch=16: $(AVL_FRAMES=240 AVL_PREPARED=1 AVL_SLAB=103,128 AVL_BATCH=16 python tools/perf_build.py 2>/dev/null | tail -1)"
