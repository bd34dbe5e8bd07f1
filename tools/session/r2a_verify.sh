#!/bin/bash
# round 2, first GPU session (1 GPU): run everything round 1 left unexecuted, then the sanitizers and the streamed-B A/B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== unverified"; AVL_UNVERIFIED=1 timeout 900 python -m pytest -m gpu tests/test_unverified_gpu.py -q 2>&1 | tail -25
echo "== p2p n=1"; timeout 300 python tools/p2p_check.py 2>&1 | tail -5
for tool in racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --log-file gpurun_out/r2a_${tool}.log python tools/sanitize_small.py > gpurun_out/r2a_${tool}.out 2>&1
  echo "$tool: rc=$? $(tail -n 1 gpurun_out/r2a_${tool}.log) | $(tail -n 1 gpurun_out/r2a_${tool}.out)"
done
echo "== bench default"; python bench.py --steps 200 --warmup 5 --no-build --no-cpu > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; cut -c1-900 gpurun_out/r2a_bench_default.json
echo "== bench streamed B"; AVL_STREAM_B=1 AVL_CTA_GROUP=2 python bench.py --steps 200 --warmup 5 --no-build --no-cpu > gpurun_out/r2a_bench_streamb.json 2> gpurun_out/r2a_bench_streamb.err; cut -c1-900 gpurun_out/r2a_bench_streamb.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
