#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== clustered (1M rows, noise 0.1)"; timeout 600 python tools/clustered_check.py --n 1048576 --noise 0.1 2>&1 | tail -5 | cut -c1-330
echo "== clustered (4M rows, noise 0.1 and 0.5)"; timeout 900 python tools/clustered_check.py 2>&1 | tail -8 | cut -c1-330
echo "== bench N=1 build lines"; python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-sustained --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['extra'].get('build',{})
print('step',d['ms_per_step'],{k:v for k,v in b.items() if k in ('hwc','hwc_batched8','hwc_f16_batched8','chw_reference_layout')})
"
