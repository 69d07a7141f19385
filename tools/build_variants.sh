#!/bin/bash
# Build occupancy variants of liblbm_b200.so for tools/run_variants2.sh (run HERE, before gpurun: nvcc
# cross-compiles; variants/ is git-ignored but travels to the GPU box).
#   tools/build_variants.sh "6 10" "7 12"      # pairs: <fp64 CTAs/SM> <fp32 CTAs/SM> of the dense step kernel
# Sparse kernel: SP64=<n> SP32=<n> tools/build_variants.sh "6 10"
set -e
cd "$(dirname "$0")/../lattice_boltzmann_method_gpu_b200/csrc"
make -s
mkdir -p ../../variants
for v in "$@"; do
  set -- $v
  extra=""
  [ -n "$SP64" ] && extra="$extra -DLBM_SP64_MINB=$SP64"
  [ -n "$SP32" ] && extra="$extra -DLBM_SP32_MINB=$SP32"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
       -DLBM_F64_MINB=$1 -DLBM_F32_MINB=$2 $extra -c lbm_step_fast.cu -o /tmp/fast_$1_$2.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/liblbm_$1_$2.so \
       lbm_geo.o lbm_step_strict.o /tmp/fast_$1_$2.o lbm_api.o lbm_voxel.o -lcudart
  echo "built variants/liblbm_$1_$2.so"
done
