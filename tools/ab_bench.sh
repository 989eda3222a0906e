#!/bin/bash
# usage: ab.sh v1 v2 ...   ("" = product lib) ; runs bench twice per variant interleaved
for rep in 1 2; do for v in "$@"; do L=; [ "$v" != "prod" ] && L=$PWD/pyspecsdr_b200/libpss_$v.so; PSS_LIB=$L python bench.py --steps 60 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][0]); k=j['roofline']['all_kernels_ms']; print('$v', round(j['ms_per_step'],4), [round(x,4) for x in k.values()], j['parity']['ok'])"; done; done
