#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): tests, smoke, bench lines, launch list of the bench command, full ncu
# captures of this library's kernels at the bench's own sizes, the per-kernel microbench table, and the DRAM
# traffic of the large-transform kernel with and without a thread-block cluster per frame.
R=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_${R}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${R}.log 2>&1
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --mode NFM --no-cpu --no-configs > gpurun_out/${R}_bench_nfm.json 2>> gpurun_out/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-steps 1 \
    > gpurun_out/${R}_bench_under_ncu.log 2>&1
# one step = psd, render, corr, 5 x (force, scan): skip the 3 warm-up steps (39 launches) and take one step
ncu --set full --clock-control none --import-source on -k regex:'psd_kernel|demod_|display_render' \
    -s 39 -c 5 -o gpurun_out/${R}_bench_kernels python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --e2e-steps 1 \
    > gpurun_out/${R}_bench_full_ncu.log 2>&1
PSD_SIZES=512,1024,2048,4096,8192,16384,32768,65536,131072,262144,524288,1048576 python tools/all_microbench.py > gpurun_out/${R}_microbench.txt 2>&1
for cl in 1 2; do
  PSS_LARGE_CL=$cl ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:psd_large -c 4 --csv --log-file gpurun_out/${R}_large_cl${cl}.csv python tools/psd_profile.py 32768 > /dev/null 2>&1
done
tail -n 2 gpurun_out/pytest_gpu_${R}.log gpurun_out/smoke_${R}.log
tail -c 400 gpurun_out/${R}_bench.json
