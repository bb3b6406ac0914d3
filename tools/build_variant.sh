#!/bin/bash
# build one experimental variant of libds2i_gpu.so:  tools/build_variant.sh <name> [nvcc -D flags...]
set -e
name=$1; shift
mkdir -p ds2i_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -DDS2I_DEV_FAST_BUILD "$@" \
  -shared -o ds2i_b200/lib/variants/libds2i_gpu_$name.so ds2i_b200/csrc/ds2i_gpu.cu
