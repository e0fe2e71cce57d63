#!/bin/bash
# round 2, GPU call 3 (one B200): GPU tests on the new defaults; increment-loop parity experiment on BASELINE's own materials/loads
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest3.log
timeout 900 python tests/experiments/exp_increment_parity.py > gpurun_out/r02_increment_parity.log 2>&1; echo "exp rc=$?"
cat gpurun_out/r02_increment_parity.log
