#!/bin/bash
# A/B timing of per-point kernel builds (tools/build_variant.py) on the evaluation protocol; one line per build
# usage: tools/tune_lanes.sh TAG N [lib ...]   (default: every tools/_build/libdmk_*.so and the in-tree library)
TAG=$1; N=${2:-128}; shift 2
LIBS="$@"; [ -z "$LIBS" ] && LIBS="solver_cpp_b200/libdamask_grid_b200.so $(ls tools/_build/libdmk_*.so 2>/dev/null)"
mkdir -p gpurun_out
for L in $LIBS; do
  DMK_LIB_PATH=$PWD/$L timeout 300 python bench.py --protocol evaluation --no-cpu --steps 5 --warmup 3 --cells $N $N $N 2>/dev/null \
    | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib':'$L','n':$N,'constitutive_ms':d['kernel_ms']['constitutive'],'lp_iter':d['inner_Lp_iterations_per_cell'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/${TAG}_tune.jsonl
done
