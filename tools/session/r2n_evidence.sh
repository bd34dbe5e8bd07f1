#!/bin/bash
# round 2 evidence (1 GPU): launch lists of the headline step and the build, ncu --set full of the screen / scatter /
# finalize kernels with their raw-page CSV exports, sanitizers (memcheck, racecheck, synccheck) on the small driver.
cd "$(dirname "$0")/../.."
T=r2
mkdir -p gpurun_out
echo "== launch list: headline step (bench.py, 3 steps)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-build --no-extra --no-sustained > /dev/null 2>&1
echo "== launch list: build"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_build.csv python tools/perf_build.py > /dev/null 2>&1
echo "== ncu full: screen kernel (main pass of a headline step)"
ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 9 -c 2 -o gpurun_out/${T}_prof_screen python tools/perf_screen.py --child 4 > /dev/null 2>&1
echo "== ncu full: topk_finalize"
ncu --set full --clock-control none --import-source on -k regex:topk_finalize -s 6 -c 1 -o gpurun_out/${T}_prof_finalize python tools/perf_screen.py --child 4 > /dev/null 2>&1
echo "== ncu full: scatter"
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 30 -c 1 -o gpurun_out/${T}_prof_scatter python tools/perf_build.py > /dev/null 2>&1
for k in screen finalize scatter; do
  ncu -i gpurun_out/${T}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_$k.csv 2>/dev/null
  ls -la gpurun_out/${T}_prof_$k.ncu-rep gpurun_out/${T}_ncu_raw_$k.csv
done
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --log-file gpurun_out/${T}_${tool}.log python tools/sanitize_small.py > gpurun_out/${T}_${tool}.out 2>&1
  echo "$tool: rc=$? $(tail -n 1 gpurun_out/${T}_${tool}.log) | $(tail -n 1 gpurun_out/${T}_${tool}.out)"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
