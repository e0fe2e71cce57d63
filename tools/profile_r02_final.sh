#!/bin/bash
# Final profile of round 2 on one B200: GPU test suite, the bench line (default: solve protocol, C3 256^3), the ncu launch list of the
# solve protocol at 128^3, `ncu --set full` of the fused NGMRES vector kernels and of the spectral kernels at 256^3.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu_final.log | cut -c1-300
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench_n1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_solve_128.csv \
    python bench.py --no-cpu --steps 1 --warmup 3 --cells 128 128 128 > gpurun_out/${TAG}_launches_solve_128.log 2>&1; echo "launch list (solve) rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches_solve_128.csv "ncu launch list: python bench.py --no-cpu --steps 1 --warmup 3 --cells 128 128 128 (solve protocol), first 3000 launches" > gpurun_out/${TAG}_launches_solve_128_summary.txt
head -14 gpurun_out/${TAG}_launches_solve_128_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:"dmk_ngmres" --launch-skip 150 -c 6 -f -o gpurun_out/${TAG}_ngmres_256 \
    python bench.py --no-cpu --steps 1 --warmup 3 > gpurun_out/${TAG}_ngmres_256.log 2>&1; echo "ngmres capture rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"dmk_fast|dmk_sum9|dmk_mech_aim" --launch-skip 40 -c 8 -f -o gpurun_out/${TAG}_fft_256 \
    python bench.py --protocol evaluation --no-cpu --steps 2 --warmup 3 > gpurun_out/${TAG}_fft_256.log 2>&1; echo "fft capture rc=$?"
ls -la gpurun_out/${TAG}_*
du -sh gpurun_out
