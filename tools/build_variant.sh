#!/bin/bash
# Builds a compile-time variant of the library next to the product build, for A/B measurements:
#   tools/build_variant.sh minb3 -DPSS_SMOOTH_MINB=3      ->  pyspecsdr_b200/libpss_minb3.so
#   PSS_LIB=$PWD/pyspecsdr_b200/libpss_minb3.so python bench.py ...
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/var_$name
for f in pyspecsdr_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude "$@" -c -o build/var_$name/$(basename $f .cu).o $f &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o pyspecsdr_b200/libpss_$name.so build/var_$name/*.o
echo built pyspecsdr_b200/libpss_$name.so
