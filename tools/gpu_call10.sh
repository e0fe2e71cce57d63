#!/bin/bash
# round 2: the other BASELINE.json configurations on one B200 through bench.py --config (solve protocol)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 420 python bench.py "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$name.json'))
    print('$name', 'value %.4g' % d['value'], 'ms/step %.1f' % d['ms_per_step'], 'evals', d['evaluations'], 'iters', d['ngmres_iterations'], 'kernel_ms', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='per'}, 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('$name: no line', e)
PY
}
run C1 --config C1 --steps 10 --warmup 3
run C2 --config C2 --steps 3 --warmup 3
run C4_128 --config C4 --steps 2 --warmup 3 --cells 128 128 128 --no-cpu
run C5_64 --config C5 --steps 2 --warmup 3 --cells 64 64 64 --no-cpu
