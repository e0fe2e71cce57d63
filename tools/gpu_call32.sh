#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_isostrain.py -q -x > gpurun_out/r02d_pytest_isostrain.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02d_pytest_isostrain.log | cut -c1-600
