#!/bin/bash
# round 2, GPU call 1 (one B200): GPU tests, the solve-protocol bench line, one ncu --set full capture of the per-point kernel
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest1.log
timeout 400 python bench.py --steps 3 --warmup 3 --cells 128 128 128 --no-cpu > gpurun_out/r02_bench_solve_128.json 2> gpurun_out/r02_bench_solve_128.err; echo "bench128 rc=$?"
cut -c1-2500 gpurun_out/r02_bench_solve_128.json
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_solve_n1.json 2> gpurun_out/r02_bench_solve_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_solve_n1.err
cut -c1-3000 gpurun_out/r02_bench_solve_n1.json
FL=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
timeout 900 ncu --set full --import-source on --clock-control none --metrics $FL -k regex:dmk_pheno_lanes -s 14 -c 4 -f -o gpurun_out/r02_lanes \
    python bench.py --protocol evaluation --no-cpu --steps 1 --warmup 3 > gpurun_out/r02_lanes_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_lanes_ncu.log | cut -c1-600
ls -la gpurun_out/r02_*
