#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== build tests"; python -m pytest tests/test_build_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== clustered"; timeout 600 python tools/clustered_check.py 2>&1 | tail -20
echo "== build e2e"; python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-sustained 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps(d['e2e'].get('build'), indent=0))
"
