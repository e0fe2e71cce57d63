#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 140 python -m pytest tests/test_gpu_constitutive.py -q > gpurun_out/r02c_pytest_constitutive_final.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02c_pytest_constitutive_final.log | cut -c1-400
