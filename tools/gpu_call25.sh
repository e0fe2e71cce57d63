#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_fullsize.py -x -q -k "not constitutive" > gpurun_out/r02c_pytest_spec.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c_pytest_spec.log | cut -c1-300
for n in 256 512; do python tools/bench_spectral.py $n 0 0 20; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_spectral_256.csv python tools/bench_spectral.py 256 0 0 3 > /dev/null 2>&1
python - <<PY
import csv, collections
d=collections.defaultdict(list)
for r in csv.reader(open('gpurun_out/r02c_launches_spectral_256.csv')):
    if len(r)>10 and r[0].isdigit(): d[r[4][:70]].append(float(r[-1])/1e6)
for k,v in d.items(): print('%-72s n=%d last %.3f ms' % (k, len(v), v[-1]))
PY
