#!/bin/bash
cd "$(dirname "$0")/../.."
AVL_DEBUG_FLAGS=128 python tools/perf_screen.py --child 30 2>&1 | grep "avl timeline" | tail -3
timeout 600 python tools/perf_screen.py --steps 20 300 --variants "pipe:"
echo "== index tests"; timeout 900 python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -2
