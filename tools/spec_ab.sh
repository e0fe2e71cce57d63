#!/bin/bash
# spectral tests + A/B of the fused z pass (pipelined / one tile per CTA) with per-kernel times from an ncu launch list
mkdir -p gpurun_out
TAG=$1
timeout 600 python -m pytest tests/test_gpu_spectral.py -q -m gpu -x 2>&1 | tail -4
for P in 1 0; do
  for N in 256 128 512; do
    echo "DMK_GAMMA_PIPE=$P N=$N: $(DMK_GAMMA_PIPE=$P python tools/bench_spectral.py $N 0 0 10)" | tee -a gpurun_out/${TAG}_spec_ab.txt
  done
done
DMK_GAMMA_PIPE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_spectral_256.csv python tools/bench_spectral.py 256 0 0 4 > /dev/null 2>&1
python - <<PY
import csv, collections
d=collections.defaultdict(list)
for r in csv.reader(open('gpurun_out/${TAG}_launches_spectral_256.csv')):
    if len(r)>10 and r[0].isdigit(): d[r[4][:70]].append(float(r[-1])/1e6)
for k,v in d.items(): print('%-72s n=%d last %.3f ms' % (k,len(v),v[-1]))
PY
