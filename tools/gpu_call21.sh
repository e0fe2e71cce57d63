#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02c_pytest_gpu.log | cut -c1-300
for mode in fused split; do
  DMK_NGMRES_VECOPS=$mode timeout 300 python bench.py --no-cpu --steps 1 --warmup 3 --cells 128 128 128 > gpurun_out/r02c_solve128_$mode.json 2> gpurun_out/r02c_solve128_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02c_solve128_$mode.json'))
print('$mode', 'value %.4g' % d['value'], 'evals', d['evaluations'], 'its', d['ngmres_iterations'], 'wall_ms %.0f' % d['wall_ms'], 'outside ms/step %.1f' % d['kernel_ms']['outside_evaluations_ms_per_step'], 'const %.2f' % d['kernel_ms']['constitutive'])
print(d['kernel_ms']['ngmres_wall_ms_per_iteration'])
PY
done
