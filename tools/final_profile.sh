#!/bin/bash
# round-end evidence: tests, bench lines, ncu launch list + full captures (copied to profiles/ by hand)
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 7 -c 1 -o gpurun_out/prof_screen_r1b python bench.py --steps 2 --warmup 3 --no-build --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:scatter_kernel -s 30 -c 1 -o gpurun_out/prof_scatter_r1b python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:argmax_rerank -s 3 -c 1 -o gpurun_out/prof_rerank_r1b python bench.py --steps 1 --warmup 3 --no-build --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
cat gpurun_out/bench_n1.json | cut -c1-3500
