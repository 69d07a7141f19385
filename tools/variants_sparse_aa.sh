#!/bin/bash
# Occupancy variants of the sparse in-place kernels (run HERE before gpurun; variants/ travels to the GPU box):
#   tools/variants_sparse_aa.sh "e64 o64 e32 o32" ...   CTAs/SM of the even/odd kernel in fp64 and fp32
# -> variants/liblbm_spaa_<e64>_<o64>_<e32>_<o32>.so; measure with LBM_B200_LIB=<that> python tools/sparse_bench.py ...
set -e
cd "$(dirname "$0")/../lattice_boltzmann_method_gpu_b200/csrc"
make -s
mkdir -p ../../variants
for v in "$@"; do
  set -- $v
  tag=$1_$2_$3_$4
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
       -DLBM_SPAA64_MINB=$1 -DLBM_SPAA64_ODD_MINB=$2 -DLBM_SPAA32_MINB=$3 -DLBM_SPAA32_ODD_MINB=$4 $EXTRA \
       -c lbm_step_fast.cu -o /tmp/fast_spaa_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/liblbm_spaa_$tag.so \
       lbm_geo.o lbm_step_strict.o /tmp/fast_spaa_$tag.o lbm_api.o lbm_voxel.o -lcudart
  echo "built variants/liblbm_spaa_$tag.so"
done
