#!/bin/bash
# multi-GPU parity record: tests/mgpu_parity.py on N ranks with both exchange paths, then a short solve-protocol bench with its
# start-up parity check.  Usage (under gpurun --gpus N): bash tools/gpu_call_mgpu.sh N [bench cells...]
set -u
N=${1:-2}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for cfg in ${DMK_MGPU_CFGS:-p2p:nccl nccl:nccl p2p:torch}; do
  ex=${cfg%%:*}; cm=${cfg##*:}
  DMK_EXCHANGE=$ex DMK_COMM=$cm DMK_PARITY_TIMEOUT=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      tests/mgpu_parity.py > gpurun_out/r02_mgpu_parity_n${N}_${ex}_${cm}.log 2>&1; echo "parity N=$N exchange=$ex collectives=$cm rc=$?"
  grep -E "spectral|solver:|update_coords|gather_slabs|collectives|MGPU_PARITY" gpurun_out/r02_mgpu_parity_n${N}_${ex}_${cm}.log
done
if [ $# -ge 3 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
      bench.py --gpus $N --steps 1 --warmup 3 --no-cpu --cells $1 $2 $3 > gpurun_out/r02_bench_solve_n${N}_small.json 2> gpurun_out/r02_bench_solve_n${N}_small.err; echo "bench N=$N rc=$?"
  tail -3 gpurun_out/r02_bench_solve_n${N}_small.err | cut -c1-400
  cut -c1-3000 gpurun_out/r02_bench_solve_n${N}_small.json
fi
if [ "${DMK_MGPU_EVAL:-1}" = "1" ]; then
  for pipe in 1 0; do
    DMK_PIPELINE=$pipe timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
        bench.py --gpus $N --protocol evaluation --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_eval_n${N}_pipe$pipe.json 2> gpurun_out/r02_bench_eval_n${N}_pipe$pipe.err; echo "bench eval N=$N pipeline=$pipe rc=$?"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_eval_n${N}_pipe$pipe.json'))
    print('N=$N pipeline=$pipe ms/eval %.2f' % d['ms_per_step'], d['kernel_ms'], 'hbm frac %.3f' % d['roofline_hbm']['frac'])
except Exception as e:
    print('no line:', e)
PY
  done
fi
