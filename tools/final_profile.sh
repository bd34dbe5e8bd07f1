#!/bin/bash
# round-end evidence (1 GPU): tests, smoke, bench lines, ncu launch list + full captures.
# Summaries are extracted here afterwards (ncu -i ... --page raw --csv) and copied to profiles/ by hand.
cd "$(dirname "$0")/.."
TAG=${1:-r1c}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 300 gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference_arm.json 2>&1
# sanitizers on tiny shapes (every kernel family; results still checked against the oracle)
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/${TAG}_${tool}.log python tools/sanitize_small.py > gpurun_out/${TAG}_${tool}.out 2>&1
  echo "$tool: $(tail -n 1 gpurun_out/${TAG}_${tool}.log)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_build.csv python tools/perf_build.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_fuse.csv python tools/perf_fuse.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_heat.csv python tools/perf_heat.py > /dev/null 2>&1
# main screen pass of a headline step: launches per step = query_prepare, sample screen, select, screen, finalize
ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 7 -c 1 -o gpurun_out/${TAG}_prof_screen python bench.py --steps 2 --warmup 3 --no-build --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:scatter_kernel -s 30 -c 1 -o gpurun_out/${TAG}_prof_scatter python tools/perf_build.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:topk_finalize -s 5 -c 1 --import-source on -o gpurun_out/${TAG}_prof_finalize python bench.py --steps 2 --warmup 3 --no-build --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:argmax_rerank -s 10 -c 1 -o gpurun_out/${TAG}_prof_rerank python bench.py --steps 1 --warmup 3 --no-build --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -14
cut -c1-600 gpurun_out/${TAG}_bench_n1.json
