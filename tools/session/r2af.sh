#!/bin/bash
cd "$(dirname "$0")/../.."
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== bench N=1 (driver form)"; python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_bench_n1.json 2> gpurun_out/r2af_bench_n1.err; tail -2 gpurun_out/r2af_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2af_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3))
b=d['extra'].get('build',{})
print({k:(round(v['frames_per_s']) if isinstance(v,dict) and 'frames_per_s' in v else None) for k,v in b.items() if k.startswith('hwc') or k.startswith('chw')})
print('errors',{k:v for k,v in d['extra'].items() if 'error' in k})
PY
