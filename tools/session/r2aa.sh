#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sharded_build_check.py 2>&1 | grep -v "^W\|Setting OMP\|^\*\*\*" | tail -1
