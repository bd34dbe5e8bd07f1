#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2t_bench_n$N.json 2> gpurun_out/r2t_bench_n$N.err; tail -2 gpurun_out/r2t_bench_n$N.err; tail -1 gpurun_out/r2t_bench_n$N.json | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d.get('parity'),'kernel',round(d['roofline']['kernel_ms'],4))
print('build_slab_sharded',d['extra'].get('build_slab_sharded'))
"
echo "== sharded build check (balanced slabs), N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sharded_build_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -1 | tee gpurun_out/r2t_build_check_n$N.json | cut -c1-420
