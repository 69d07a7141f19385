#!/bin/bash
# round-2 GPU call 5: sparse in-place storage with its own numbering -- tests, self-check, throughput, occupancy variants
cd "$(dirname "$0")/.."
O=gpurun_out/r2c5; mkdir -p $O
timeout 1500 python -m pytest tests/test_sparse_aa_gpu.py tests/test_group_gpu.py tests/test_reinit_gpu.py tests/test_sparse_gpu.py tests/test_edge_cases_gpu.py tests/test_checkpoint_gpu.py -m gpu -q -p no:cacheprovider > $O/pytest_sparse.log 2>&1; echo "pytest exit $?" >> $O/pytest_sparse.log
tail -15 $O/pytest_sparse.log
timeout 900 python tools/selfcheck.py > $O/selfcheck.log 2>&1; cat $O/selfcheck.log
for prec in f64 f32; do
  timeout 600 python tools/sparse_bench.py --n 512 --steps 50 --precision $prec --only sparse_aa > $O/sparse_${prec}_default.json 2> $O/err.txt
  python -c "import json;d=json.load(open('$O/sparse_${prec}_default.json'))['sparse_aa'];print('default $prec', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4), d['device_GB'], d['nlattice'])"
  for v in 6_5_10_8 7_5_12_9 5_4_8_7; do
    LBM_B200_LIB=$PWD/variants/liblbm_spaa_$v.so timeout 600 python tools/sparse_bench.py --n 512 --steps 50 --precision $prec --only sparse_aa > $O/sparse_${prec}_$v.json 2>> $O/err.txt
    python -c "import json;d=json.load(open('$O/sparse_${prec}_$v.json'))['sparse_aa'];print('$v $prec', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4))"
  done
done
tail -3 $O/err.txt
