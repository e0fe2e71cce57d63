#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_constitutive.py -x -q -k "kinehardening or other_state or tangent" > gpurun_out/r02c_pytest_kine.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02c_pytest_kine.log | cut -c1-400
