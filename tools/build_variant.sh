#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]: builds variants/NAME.so (tuning experiments, loaded via CHS_B200_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC "$@" \
     -o variants/$name.so chsimpy_b200/csrc/chs_api.cu chsimpy_b200/csrc/chs_ll.cu
echo "built variants/$name.so"
