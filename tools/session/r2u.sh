#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index tests"; timeout 900 python -m pytest tests/test_index_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/perf_screen.py --steps 20 300 --variants "pipe:"
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-build --no-cpu --no-extra 2> gpurun_out/r2u_bench.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3),'stepfrac',round(r['step_frac_vs_burst_peak'],3))
print('sustained',r.get('sustained'))
"; tail -3 gpurun_out/r2u_bench.err
echo "== bench N=1 no pipeline"; AVL_BENCH_NO_PIPELINE=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-build --no-cpu --no-extra --no-sustained 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']))
"
