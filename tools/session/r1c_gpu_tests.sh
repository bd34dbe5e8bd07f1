#!/bin/bash
# GPU parity suite + smoke (1 GPU)
cd "$(dirname "$0")/../.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/sharded_build_check.py > gpurun_out/sharded_build_n1.json 2> gpurun_out/sharded_build_n1.err; tail -c 300 gpurun_out/sharded_build_n1.err; cat gpurun_out/sharded_build_n1.json
