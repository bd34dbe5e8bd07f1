#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2f_launches_step.csv python tools/perf_screen.py --child 4 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2f_launches_step.csv')) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Metric Unit, Metric Value
last=rows[-40:]
for r in last: print(r[4][:60].ljust(60), r[8], r[-1], r[-2])
PY
