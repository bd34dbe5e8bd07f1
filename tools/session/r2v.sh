#!/bin/bash
cd "$(dirname "$0")/../.."
AVL_DEBUG_FLAGS=128 python tools/perf_screen.py --child 30 2>&1 | grep "avl timeline" | tail -6
