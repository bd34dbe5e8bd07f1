#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
echo "== build tests"; python -m pytest tests/test_build_gpu.py tests/test_handoff_p2p_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== bench N=1 build lines"; python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-sustained --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['extra'].get('build',{})
print({k:v for k,v in b.items() if k in ('hwc','hwc_batched8','hwc_f16_batched8','chw_reference_layout','hwc_f16_error')})
print(d['roofline'].get('build_scatter'))
"
echo "== sharded build check (balanced slabs), N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sharded_build_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -1 | cut -c1-330
