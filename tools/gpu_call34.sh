#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02e_final.log
( timeout 40 python -m pytest -q -x tests/test_gpu_isostrain.py tests/test_gpu_consumer.py ; echo "pytest rc=$?" ) > $L 2>&1
( timeout 25 python bench.py --config C1 --steps 3 --warmup 3 > gpurun_out/r02e_bench_C1.json 2>gpurun_out/r02e_bench_C1.err ; echo "bench rc=$?" ) >> $L 2>&1
tail -25 $L | cut -c1-400; head -c 600 gpurun_out/r02e_bench_C1.json
