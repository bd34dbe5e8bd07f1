#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench N=1 (driver form)"; ( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err ) 2>&1 | grep real; tail -3 gpurun_out/r2o_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3))
print('build e2e',json.dumps(d['e2e'].get('build')))
print('dropin',json.dumps(d['extra'].get('dropin')), d['extra'].get('dropin_error'))
print('cpu config3', d['cpu_baseline'].get('config3'))
PY
bash tools/session/r2n_evidence.sh
