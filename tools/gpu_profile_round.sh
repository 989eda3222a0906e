#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): tests, smoke, bench line, launch list of the bench command and
# full ncu captures of this library's kernels at the bench's own sizes.
R=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_${R}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${R}.log 2>&1
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
python bench.py --mode NFM --no-cpu > gpurun_out/bench_nfm_${R}.json 2>> gpurun_out/bench_${R}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'psd_kernel|demod_decim_kernel|display_render' \
    -s 9 -c 3 -o gpurun_out/bench_kernels_${R} python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 \
    > gpurun_out/bench_full_ncu_${R}.log 2>&1
tail -n 2 gpurun_out/pytest_gpu_${R}.log gpurun_out/smoke_${R}.log
tail -c 600 gpurun_out/bench_${R}.json
