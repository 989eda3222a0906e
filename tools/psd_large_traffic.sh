#!/bin/bash
# DRAM traffic per launch of the large-transform PSD kernels (run on the GPU box).
for N in 16384 32768 65536; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'psd_large|psd_kernel|colfft|row_epilogue' -s 2 -c 2 --csv python tools/psd_profile.py $N 2>/dev/null \
      | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
for r in rows[1:]:
    print($N, r[4][:40], r[-3], r[-2], r[-1])
"
done
