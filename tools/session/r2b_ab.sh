#!/bin/bash
# round 2, session b: tile-major operand copy + L2 prefetch warp + deferred candidate flush: parity, then A/B at burst
# (20 steps) and power-capped (300 steps) clocks, then the triage flags at burst clocks.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index tests (tiled default)"; python -m pytest tests/test_index_gpu.py tests/test_unverified_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== index tests (AVL_TILED=0)"; AVL_TILED=0 python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== index tests (prefetch 2)"; AVL_PREFETCH_TILES=2 python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2b_ab.json --variants \
  "rowmajor:AVL_TILED=0" "tiled:" "tiled_pf1:AVL_PREFETCH_TILES=1" "tiled_pf2:AVL_PREFETCH_TILES=2" "tiled_pf4:AVL_PREFETCH_TILES=4" \
  "ts:AVL_TS=1" "ts_pf0_rowmajor:AVL_TS=1,AVL_TILED=0"
echo "== triage at burst clocks (tiled, no prefetch)"
python tools/perf_screen.py --steps 20 --out gpurun_out/r2b_triage.json --variants \
  "f4_nodrain:AVL_DEBUG_FLAGS=4" "f16_noemit:AVL_DEBUG_FLAGS=16" "f32_ldonly:AVL_DEBUG_FLAGS=32" "f8_fullcmp:AVL_DEBUG_FLAGS=8" \
  "f1_nomma:AVL_DEBUG_FLAGS=1" "f5_tmaonly:AVL_DEBUG_FLAGS=5" "f6_mmaonly:AVL_DEBUG_FLAGS=6" "f2_noload:AVL_DEBUG_FLAGS=2" \
  "f4_pf2:AVL_DEBUG_FLAGS=4,AVL_PREFETCH_TILES=2" "f5_pf2:AVL_DEBUG_FLAGS=5,AVL_PREFETCH_TILES=2"
