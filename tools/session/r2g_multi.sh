#!/bin/bash
# multi-GPU proof: config 5 parity for both exchanges + strong scaling line; p2p micro check; weak-scaling bench lines
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
echo "== gpu suite (index + p2p)"; python -m pytest tests/test_index_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -2
if [ "$N" = "2" ]; then echo "== config 5, N=1"; timeout 600 python tools/sharded_index_check.py 2>&1 | tail -2; fi
echo "== p2p_check N=$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py 2>&1 | grep -v "^W\|Setting OMP" | tail -3
echo "== config 5, N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/sharded_index_check.py 2>&1 | grep -v "^W\|Setting OMP" | tail -3
