#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest7.log
