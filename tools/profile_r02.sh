#!/bin/bash
# Round-2 profile on one B200 (run under gpurun): (1) the bench line itself, (2) ncu launch lists (solve protocol at 128^3, first
# 3000 launches; evaluation protocol at 256^3), (3) `ncu --set full` of ONE launch of the per-point kernel at 256^3 with source
# counters and FP64 instruction counts, (4) `ncu --set full` of the FFT / Gamma kernels of one evaluation.  Outputs in gpurun_out/
# (kept under 64 MiB in total: larger results are not copied back).
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_solve_128.csv \
    python bench.py --no-cpu --steps 1 --warmup 3 --cells 128 128 128 > gpurun_out/${TAG}_launches_solve_128.log 2>&1; echo "launch list (solve) rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_eval_256.csv \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_launches_eval_256.log 2>&1; echo "launch list (evaluation) rc=$?"
FL=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
# main launches of the per-point kernel alternate with the (empty) launch for the given-up list: even indices are main launches
timeout 900 ncu --set full --import-source on --clock-control none --metrics $FL -k regex:dmk_pheno_lanes --launch-skip 10 -c 1 -f -o gpurun_out/${TAG}_lanes_256 \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_lanes_256.log 2>&1; echo "lanes capture rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"dmk_fast|dmk_sum9|dmk_mech_aim" --launch-skip 40 -c 8 -f -o gpurun_out/${TAG}_fft_256 \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_fft_256.log 2>&1; echo "fft capture rc=$?"
ls -la gpurun_out/${TAG}_*
du -sh gpurun_out
