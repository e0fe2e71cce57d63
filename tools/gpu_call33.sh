#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02d_regression.log
( timeout 70 python -m pytest -q -x tests/test_gpu_consumer.py "tests/test_gpu_solver.py::test_fortran_named_dropin_symbols" "tests/test_gpu_solver.py::test_form_residual_host_roundtrip" tests/test_gpu_constitutive.py tests/test_gpu_isostrain.py ; echo "pytest rc=$?" ) > $L 2>&1
( timeout 25 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" ; echo "smoke rc=$?" ) >> $L 2>&1
tail -30 $L | cut -c1-500
