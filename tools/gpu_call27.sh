#!/bin/bash
# C5 at its own size on one GPU (dislotwin in the reference-order kernel), solve protocol, 2 increments
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 > gpurun_out/r02_bench_C5_256.json 2> gpurun_out/r02_bench_C5_256.err; echo "C5 rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_C5_256.json'))
    print('C5', 'value %.4g' % d['value'], 'evals', d['evaluations'], 'its', d['ngmres_iterations'], 'wall_ms %.0f' % d['wall_ms'], 'const %.3f' % d['kernel_ms']['constitutive'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'e2e', d['e2e']['value'])
except Exception as e:
    print('no line', e); print(open('gpurun_out/r02_bench_C5_256.err').read()[-1500:])
PY
