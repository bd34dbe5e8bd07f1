#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
echo "== config 5, N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/sharded_index_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -2
echo "== sharded build check, N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sharded_build_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -2
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err; tail -2 gpurun_out/r2l_bench_n$N.err; tail -1 gpurun_out/r2l_bench_n$N.json | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d.get('parity'),'kernel',round(d['roofline']['kernel_ms'],4))
print('build_slab_sharded',d['extra'].get('build_slab_sharded'))
"
echo "== bench N=$N reference arm"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2l_ref_n$N.json 2> /dev/null ) 2>&1 | grep real; tail -1 gpurun_out/r2l_ref_n$N.json | cut -c1-300
