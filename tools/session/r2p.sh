#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/perf_screen.py --steps 20 --variants "final:" 
echo "== bench N=1 (driver form)"; ( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err ) 2>&1 | grep real; tail -3 gpurun_out/r2p_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3),'stepfrac',round(r['step_frac_vs_burst_peak'],3),'traffic',r['traffic'])
print('sustained',json.dumps(r.get('sustained')))
print('build roofline',json.dumps(r.get('build_scatter')))
print('build e2e',json.dumps(d['e2e'].get('build')))
b=d['extra'].get('build',{})
print({k:v for k,v in b.items() if k in ('hwc','hwc_batched8','hwc_f16_batched8','chw_reference_layout','hwc_f16_error')})
print('errors',{k:v for k,v in d['extra'].items() if 'error' in k})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2p_launches_step.csv python tools/perf_screen.py --child 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2p_launches_step.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-7:]: print(r[4][:50].ljust(50), r[8], r[-1])
PY
