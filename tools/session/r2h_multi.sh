#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
echo "== config 5, N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/sharded_index_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -2
echo "== bench N=1"; python bench.py --gpus 1 --steps 20 --warmup 5 --no-build --no-cpu --no-extra > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -3 gpurun_out/r2h_bench_n1.err; cut -c1-700 gpurun_out/r2h_bench_n1.json
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-build > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err; tail -3 gpurun_out/r2h_bench_n$N.err; tail -1 gpurun_out/r2h_bench_n$N.json | cut -c1-700
echo "== bench N=$N nccl exchange"; AVL_P2P_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --no-build > gpurun_out/r2h_bench_n${N}_nccl.json 2> /dev/null; tail -1 gpurun_out/r2h_bench_n${N}_nccl.json | cut -c1-400
