#!/bin/bash
# final verification of round 2 (1 GPU): suite, smoke, bench (driver form), reference arm, evidence refresh
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench N=1 (driver form)"; ( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err ) 2>&1 | grep real; tail -3 gpurun_out/r2x_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3),'stepfrac',round(r['step_frac_vs_burst_peak'],3),'traffic',r['traffic'], 'clocks', d['clocks'])
print('sustained',json.dumps(r.get('sustained')))
print('errors',{k:v for k,v in d['extra'].items() if 'error' in k})
print('top-level keys',list(d.keys()))
PY
bash tools/session/r2n_evidence.sh 2>&1 | tail -12
