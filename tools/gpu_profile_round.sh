#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): launch list of the bench command + full captures of our kernels.
set -x
R=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --blocks 2048 \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'psd_kernel|demod_decim_kernel|display_render' \
    -s 9 -c 3 -o gpurun_out/bench_kernels_${R} python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --blocks 2048 \
    > gpurun_out/bench_full_ncu_${R}.log 2>&1
tail -2 gpurun_out/bench_full_ncu_${R}.log
