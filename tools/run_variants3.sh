#!/bin/bash
# sparse-kernel occupancy variants, then one ncu --set full capture of the default dense kernel
cd /root/repo; mkdir -p gpurun_out
LIB=lattice_boltzmann_method_gpu_b200/liblbm_b200.so
cp $LIB /tmp/lib_default.so
for v in default sp_6_10; do
  if [ $v = default ]; then cp /tmp/lib_default.so $LIB; else cp variants/liblbm_$v.so $LIB; fi
  for pr in f64 f32; do
    echo -n "$v $pr "; python tools/sparse_bench.py --steps 30 --precision $pr --only sparse_ab 2>&1 | grep -E '"mlups"|algorithmic' | tr -d '\n'; echo
  done
done
cp /tmp/lib_default.so $LIB
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_short.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_dense -s 4 -c 2 -o gpurun_out/r01_aa_f64_v2 -f \
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
