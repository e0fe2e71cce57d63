#!/bin/bash
# ncu --set full (+ source counters) of ONE launch of the per-point kernel on the evaluation protocol: tools/ncu_lanes.sh TAG N lib
TAG=$1; N=$2; L=$3
DMK_LIB_PATH=$PWD/$L timeout 600 ncu --set full --import-source on --clock-control none -k regex:dmk_pheno_lanes --launch-skip 6 -c 1 -f \
  -o gpurun_out/${TAG} python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 --cells $N $N $N > gpurun_out/${TAG}.log 2>&1
echo "ncu $TAG rc=$?"; ls -la gpurun_out/${TAG}.ncu-rep
