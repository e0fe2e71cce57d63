#!/bin/bash
# round 2, GPU call 4 (one B200): occupancy sweep of the lane-persistent kernel (CTAs per SM it is compiled for), GPU tests
set -u
mkdir -p gpurun_out
for occ in 4 5 6; do
  DMK_PL_OCC=$occ timeout 300 python bench.py --protocol evaluation --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_occ_$occ.json 2> gpurun_out/r02_occ_$occ.err; echo "occ $occ rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_occ_$occ.json'))
print('occ $occ', 'ms/eval %.2f' % d['ms_per_step'], d['kernel_ms'], 'Lp/cell %.2f' % d['inner_Lp_iterations_per_cell'], 'fp64 frac %.3f' % d['roofline']['frac'])
PY
done
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest4.log
