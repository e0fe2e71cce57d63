#!/bin/bash
# Round profile of bench.py's step on one B200 (run under gpurun): launch list of the whole command, then one
# `ncu --set full` capture of every kernel of ONE residual evaluation of the timed state.  Outputs in gpurun_out/.
set -u
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_bench_256.csv \
    python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_launches_bench_256.log 2>&1
SKIP=$(python - <<PY
import csv
ids=[]
for r in csv.reader(open('gpurun_out/${TAG}_launches_bench_256.csv')):
    if len(r) > 10 and r[0].isdigit() and 'dmk_' in r[4] and 'dfma' not in r[4]:
        ids.append(r[4])
# start of the 6th evaluation = 6th launch of the lane kernel (init, 2 Taylor, 2 relax come first)
n = 0
for k, name in enumerate(ids):
    if 'lanes' in name:
        n += 1
        if n == 7:
            print(k); break
PY
)
echo "skip=$SKIP"
ncu --set full --import-source on --clock-control none -k regex:dmk_ --launch-skip $SKIP -c 12 -o gpurun_out/${TAG}_step_full \
    python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_step_full.log 2>&1
tail -2 gpurun_out/${TAG}_step_full.log | cut -c1-300
