#!/bin/bash
# round-2 GPU call 4: self-check rerun, remaining sparse tests, ncu --set full of the sparse in-place kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r2c4; mkdir -p $O
timeout 900 python tools/selfcheck.py > $O/selfcheck.log 2>&1; cat $O/selfcheck.log
timeout 1200 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_outputs.py -m gpu -q -p no:cacheprovider > $O/pytest_sparse.log 2>&1; echo "pytest exit $?" >> $O/pytest_sparse.log
tail -8 $O/pytest_sparse.log
CMD="python tools/sparse_bench.py --n 512 --steps 4 --precision f64 --only sparse_aa"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa -s 5 -c 2 -o $O/sparse_aa_f64 $CMD > $O/ncu_f64.log 2>&1
tail -3 $O/ncu_f64.log
CMD="python tools/sparse_bench.py --n 512 --steps 4 --precision f32 --only sparse_aa"
$CMD > $O/plain32.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa -s 5 -c 2 -o $O/sparse_aa_f32 $CMD > $O/ncu_f32.log 2>&1
tail -3 $O/ncu_f32.log
ls -la $O
