#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "cu_bench_iterate" > gpurun_out/r02c_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02c_pytest_full.log | cut -c1-300
for v in 0 1 2 3 4; do echo "DMK_GZ_VARIANT=$v"; DMK_GZ_VARIANT=$v python tools/bench_spectral.py 256 0 0 20; done
DMK_GZ_VARIANT=1 timeout 300 python -m pytest tests/test_gpu_spectral.py -x -q -k "256 or thin" 2>&1 | tail -2
