#!/bin/bash
cd "$(dirname "$0")/../.."
N=2
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== bench N=1 (driver form)"; python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ad_bench_n1.json 2> gpurun_out/r2ad_bench_n1.err; tail -2 gpurun_out/r2ad_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ad_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'kernel',round(r['kernel_ms'],4),'frac',round(r['frac'],3))
print('build roofline',json.dumps(r.get('build_scatter'))[:400])
b=d['extra'].get('build',{})
print({k:(round(v['frames_per_s']) if isinstance(v,dict) and 'frames_per_s' in v else v) for k,v in b.items() if k.startswith('hwc') or k.startswith('chw')})
print('errors',{k:v for k,v in d['extra'].items() if 'error' in k})
PY
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2ad_bench_n$N.json 2> gpurun_out/r2ad_bench_n$N.err; tail -2 gpurun_out/r2ad_bench_n$N.err; tail -1 gpurun_out/r2ad_bench_n$N.json | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d.get('parity'),'kernel',round(d['roofline']['kernel_ms'],4))
print('build_slab_sharded',d['extra'].get('build_slab_sharded'))
"
echo "== reference arm N=$N (torchrun)"; ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 | cut -c1-260 ) 2>&1 | grep -v "^$\|user\|sys"
