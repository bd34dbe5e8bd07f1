#!/bin/bash
# round 2, session c: warp roles reordered (single-thread roles on the highest warp ids), candidate columns queued in
# registers (no shared-memory ring) -> 6 pipeline stages at 256 resident queries.  Parity first, then A/B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index + dropin tests"; python -m pytest tests/test_index_gpu.py tests/test_dropin_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2c_ab.json --variants \
  "stages6:" "stages5:AVL_MAX_STAGES=5" "stages4:AVL_MAX_STAGES=4" "rowmajor6:AVL_TILED=0"
python tools/perf_screen.py --steps 20 --out gpurun_out/r2c_triage.json --variants \
  "f4_nodrain:AVL_DEBUG_FLAGS=4" "f16_noemit:AVL_DEBUG_FLAGS=16" "f8_fullcmp:AVL_DEBUG_FLAGS=8" \
  "f5_tmaonly:AVL_DEBUG_FLAGS=5" "f6_mmaonly:AVL_DEBUG_FLAGS=6" "f4_s5:AVL_DEBUG_FLAGS=4,AVL_MAX_STAGES=5"
AVL_DEBUG_FLAGS=64 python tools/perf_screen.py --child 20 2>&1 | grep "avl clock" | tail -3
AVL_DEBUG_FLAGS=70 python tools/perf_screen.py --child 20 2>&1 | grep "avl clock" | tail -3
AVL_DEBUG_FLAGS=64 python tools/perf_screen.py --child 300 2>&1 | grep "avl clock" | tail -2
