#!/bin/bash
# Round-2 profile on one B200 (run under gpurun): (1) the bench line itself, (2) the ncu launch list of the solve protocol (same
# command at 128^3, one increment: ~10 k launches), (3) one `ncu --set full` capture of every kernel of ONE residual evaluation
# at 256^3 (evaluation protocol), with the FP64 instruction counts of the per-point kernel.  Outputs in gpurun_out/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/${TAG}_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${TAG}_launches_solve_128.csv \
    python bench.py --no-cpu --steps 1 --warmup 3 --cells 128 128 128 > gpurun_out/${TAG}_launches_solve_128.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/${TAG}_launches_solve_128.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_eval_256.csv \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_launches_eval_256.log 2>&1
SKIP=$(python - <<PY
import csv
big = []
k = -1
for r in csv.reader(open('gpurun_out/${TAG}_launches_eval_256.csv')):
    if len(r) > 10 and r[0].isdigit() and 'dmk_' in r[4] and 'dfma' not in r[4]:
        k += 1
        if 'pheno_lanes' in r[4] and float(r[-1]) > 5.0e7:          # a per-point launch doing a whole evaluation (> 50 ms)
            big.append(k)
print(big[-2] if len(big) >= 2 else 0)                                # the first of the two timed evaluations
PY
)
echo "skip=$SKIP"
FL=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
timeout 900 ncu --set full --import-source on --clock-control none --metrics $FL -k regex:dmk_ --launch-skip ${SKIP:-0} -c 16 -f -o gpurun_out/${TAG}_step_full \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_step_full.log 2>&1; echo "full capture rc=$?"
tail -2 gpurun_out/${TAG}_step_full.log | cut -c1-300
ls -la gpurun_out/${TAG}_*
