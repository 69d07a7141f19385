#!/bin/bash
# round-2 GPU call 3: sparse in-place storage -- parity tests, self-check build, vessel-bundle throughput
cd "$(dirname "$0")/.."
O=gpurun_out/r2c3; mkdir -p $O
timeout 1200 python -m pytest tests/test_sparse_aa_gpu.py tests/test_group_gpu.py tests/test_reinit_gpu.py tests/test_aa_gpu.py tests/test_sparse_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_sparse.log 2>&1; echo "pytest exit $?" >> $O/pytest_sparse.log
tail -30 $O/pytest_sparse.log
timeout 900 python tools/selfcheck.py > $O/selfcheck.log 2>&1; cat $O/selfcheck.log
timeout 600 python tools/sparse_bench.py --n 512 --steps 50 --precision f64 --only sparse_ab,sparse_aa > $O/sparse_f64.json 2> $O/sparse_f64.err; cat $O/sparse_f64.json; tail -3 $O/sparse_f64.err
timeout 600 python tools/sparse_bench.py --n 512 --steps 50 --precision f32 --only sparse_ab,sparse_aa > $O/sparse_f32.json 2> $O/sparse_f32.err; cat $O/sparse_f32.json; tail -3 $O/sparse_f32.err
