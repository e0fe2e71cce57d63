#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py -x -q -k "Steel_kine" > gpurun_out/r02c_pytest_kine2.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02c_pytest_kine2.log | cut -c1-400
