#!/bin/bash
# round-2 GPU call 8: where does a 64^3 step go?  ncu of the one-launch kernels and of the persistent kernel
cd "$(dirname "$0")/.."
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_variants.py tests/test_reference_outputs.py -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for prec in f64 f32; do python tools/sparse_bench.py --n 512 --steps 50 --precision $prec --only sparse_aa | python -c "import json,sys;d=json.load(sys.stdin)['sparse_aa'];print('sparse_aa $prec', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4))"; done
for p in 0 1; do for cs in ldc pos bif; do python tools/small_case.py --case $cs --persistent $p --steps 200 --calls 2 | tail -1; done; done
CMD="python tools/small_case.py --case ldc --persistent 0 --steps 6 --calls 1"
$CMD > $O/plain0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa -s 8 -c 2 -o $O/small_ldc_f32 $CMD > $O/ncu0.log 2>&1; tail -1 $O/ncu0.log
CMD="python tools/small_case.py --case ldc --persistent 1 --steps 20 --calls 1"
$CMD > $O/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa_persist -s 1 -c 1 -o $O/small_ldc_f32_persist $CMD > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
