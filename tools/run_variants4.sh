#!/bin/bash
# occupancy variants of the sparse in-place kernels after the link lists (tools/variants_sparse_aa.sh "e64 o64 e32 o32" ...)
cd "$(dirname "$0")/.."
for v in 6_5_10_9 6_5_11_9 6_5_9_9; do
  if [ $v = default ]; then unset LBM_B200_LIB; else export LBM_B200_LIB=$PWD/variants/liblbm_spaa_$v.so; fi
  for pr in f32; do
    python tools/sparse_bench.py --only sparse_aa --precision $pr --steps 40 | python -c "import json,sys; d=json.load(sys.stdin)['sparse_aa']; print('$v', '$pr', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4))"
  done
done
