#!/bin/bash
# bench lines of the other BASELINE.json configurations on one GPU (C5 at 64^3: the dislotwin law runs in the reference-order kernel)
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --config C1 --steps 3 --warmup 3 > gpurun_out/r02_bench_C1.json 2> gpurun_out/r02_bench_C1.err; echo "C1 rc=$?"
timeout 400 python bench.py --config C2 --steps 3 --warmup 3 > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; echo "C2 rc=$?"
timeout 600 python bench.py --config C5 --cells 64 64 64 --steps 3 --warmup 3 > gpurun_out/r02_bench_C5_64.json 2> gpurun_out/r02_bench_C5_64.err; echo "C5 rc=$?"
for f in C1 C2 C5_64; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$f.json'))
    print('$f', 'value %.4g' % d['value'], 'evals', d['evaluations'], 'its', d['ngmres_iterations'], 'wall_ms %.0f' % d['wall_ms'], 'const %.3f' % d['kernel_ms']['constitutive'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'e2e', d['e2e']['value'])
except Exception as e:
    print('$f', 'no line', e); print(open('gpurun_out/r02_bench_$f.err').read()[-1500:])
PY
done
