#!/bin/bash
# build a variant of libpagk_cuda.so for kernel A/B runs: tools/build_variant.sh <name> <extra nvcc flags...>  ->  tools/libpagk_<name>.so
# (run with PAGK_LIB=tools/libpagk_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=pixel_aware_gyro_aided_klt_feature_tracker_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 -shared -cudart static "$@" \
  -o tools/libpagk_$name.so $C/pagk_kernels.cu $C/pagk_lk_lanes.cu $C/pagk_api.cu
echo tools/libpagk_$name.so
