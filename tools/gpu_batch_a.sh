#!/bin/bash
mkdir -p gpurun_out
bash tools/tune_lanes.sh r02m 256
FL=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
timeout 900 ncu --set full --import-source on --clock-control none --metrics $FL -k regex:dmk_pheno_lanes --launch-skip 14 -c 1 -f -o gpurun_out/r02_lanes_256 \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/r02_lanes_256.log 2>&1; echo "lanes capture rc=$?"
for C in C1 C2; do timeout 300 python bench.py --config $C --steps 10 --warmup 3 > gpurun_out/r02_bench_$C.json 2> gpurun_out/r02_bench_$C.err; echo "$C rc=$?"; done
timeout 600 python bench.py --config C4 --cells 128 128 128 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_C4_128.json 2> gpurun_out/r02_bench_C4_128.err; echo "C4 rc=$?"
timeout 600 python bench.py --config C5 --cells 64 64 64 --steps 4 --warmup 3 --no-cpu > gpurun_out/r02_bench_C5_64.json 2> gpurun_out/r02_bench_C5_64.err; echo "C5 rc=$?"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "cu_bench" 2>&1 | tail -5
