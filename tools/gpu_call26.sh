#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py -x -q > gpurun_out/r02c_pytest_spec.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c_pytest_spec.log | cut -c1-300
for v in 0 16 163 5; do echo "DMK_XC2R_PLAN=$v"; DMK_XC2R_PLAN=$v python tools/bench_spectral.py 256 0 0 20; done
python tools/bench_spectral.py 512 0 0 10
python tools/bench_spectral.py 128 0 0 20
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_spectral_256.csv python tools/bench_spectral.py 256 0 0 3 > /dev/null 2>&1
python - <<PY
import csv, collections
d=collections.defaultdict(list)
for r in csv.reader(open('gpurun_out/r02c_launches_spectral_256.csv')):
    if len(r)>10 and r[0].isdigit(): d[r[4][:70]].append(float(r[-1])/1e6)
for k,v in d.items(): print('%-72s n=%d last %.3f ms' % (k, len(v), v[-1]))
PY
