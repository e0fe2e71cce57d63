#!/bin/bash
# round 2, GPU call 2 (one B200): GPU tests; mass-deferral variants on a virgin evaluation; solve-protocol bench at 128^3 and 256^3
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest2.log
for m in nested machine lanes; do
  DMK_PHENO_MASS=$m timeout 300 python tools/bench_constitutive.py 128 Cu_pheno 2 > gpurun_out/r02_mass_$m.json 2> gpurun_out/r02_mass_$m.err; echo "mass $m rc=$?"
  cut -c1-700 gpurun_out/r02_mass_$m.json
done
timeout 400 python bench.py --steps 3 --warmup 3 --cells 128 128 128 --no-cpu > gpurun_out/r02_bench_solve_128.json 2> gpurun_out/r02_bench_solve_128.err; echo "bench128 rc=$?"
cut -c1-2500 gpurun_out/r02_bench_solve_128.json
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_solve_n1.json 2> gpurun_out/r02_bench_solve_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_solve_n1.err
cut -c1-4000 gpurun_out/r02_bench_solve_n1.json
