#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solver.py -q -k polarisation > gpurun_out/r02_pytest_pol.log 2>&1; echo "pytest pol rc=$?"
tail -40 gpurun_out/r02_pytest_pol.log | cut -c1-300
