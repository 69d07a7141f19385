#!/bin/bash
# occupancy variants of the step kernel: build them first with tools/build_variants.sh "6 10" "7 12" (variants/liblbm_<f64 CTAs per SM>_<f32 CTAs per SM>.so)
cd /root/repo
LIB=lattice_boltzmann_method_gpu_b200/liblbm_b200.so
cp $LIB /tmp/lib_default.so
for v in default 6_10 7_12; do
  if [ $v = default ]; then cp /tmp/lib_default.so $LIB; else cp variants/liblbm_$v.so $LIB; fi
  for cfg in "f64 aa" "f64 ab" "f32 aa" "f32 ab"; do set -- $cfg
    python bench.py --steps 50 --no-cpu --no-e2e --precision $1 --storage $2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$1', '$2', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
  done
done
cp /tmp/lib_default.so $LIB
