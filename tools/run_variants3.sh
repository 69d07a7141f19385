#!/bin/bash
# occupancy variants of the dense in-place step kernel after the speculative loads became real
# (tools/build_variants.sh "5 6" "7 7" "6 8" "6 12" first: <fp64 CTAs/SM> <fp32 CTAs/SM>)
cd "$(dirname "$0")/.."
for v in default 5_6 7_7 6_8 6_12; do
  if [ $v = default ]; then unset LBM_B200_LIB; else export LBM_B200_LIB=$PWD/variants/liblbm_$v.so; fi
  for pr in f64 f32; do
    python bench.py --steps 50 --no-cpu --no-e2e --no-parity --precision $pr --storage aa | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', '$pr', 'aa', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
  done
done
