#!/bin/bash
# timing of the in-tree library on the evaluation protocol under different environment settings: tools/tune_env.sh TAG N "ENV=.." ...
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
for E in "$@"; do
  env $E timeout 300 python bench.py --protocol evaluation --no-cpu --steps 5 --warmup 3 --cells $N $N $N 2>/dev/null \
    | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'env':'$E','n':$N,'constitutive_ms':d['kernel_ms']['constitutive'],'lp_iter':d['inner_Lp_iterations_per_cell']}))" | tee -a gpurun_out/${TAG}_tune.jsonl
done
