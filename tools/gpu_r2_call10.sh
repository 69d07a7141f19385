#!/bin/bash
# round-2 GPU call 10: precomputed boundary link lists in the sparse in-place kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r2c10; mkdir -p $O
timeout 1200 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_variants.py tests/test_drivers_gpu.py tests/test_group_gpu.py tests/test_reinit_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for prec in f32 f64; do for p in 0 1; do for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision $prec --persistent $p --steps 400 --calls 2 | tail -1; done; done; done 2>&1 | tee $O/small_case.txt
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
for prec in f64 f32; do python tools/sparse_bench.py --n 512 --steps 50 --precision $prec --only sparse_aa | python -c "import json,sys;d=json.load(sys.stdin)['sparse_aa'];print('sparse_aa $prec', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4))"; done
timeout 600 python tools/selfcheck.py > $O/selfcheck.log 2>&1; cat $O/selfcheck.log
