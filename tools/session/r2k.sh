#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== build tests, dense H2D"; AVL_BUILD_DENSE_H2D=1 python -m pytest tests/test_build_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench N=1 (driver form)"; ( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err ) 2>&1 | grep real; tail -3 gpurun_out/r2k_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3),'traffic',r['traffic'])
print('sustained',r.get('sustained'))
print('cpu',d.get('cpu_baseline'))
print('build roofline',r.get('build_scatter'))
print('build e2e',d['e2e'].get('build'))
print('extra keys',list(d['extra'].keys()))
for k in ('config2_1M_x512_q64','config2_1M_x512_q64_f16_operands','config3_fusion_1M_512+1024_32pairs','build_error','config2_error','config3_error'):
    if k in d['extra']: print(k, d['extra'][k])
b=d['extra'].get('build',{})
print({k:v for k,v in b.items() if k in ('hwc','hwc_batched8','chw_reference_layout')})
PY
echo "== reference arm"; ( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err ) 2>&1 | grep real; cut -c1-1200 gpurun_out/r2k_bench_ref.json
