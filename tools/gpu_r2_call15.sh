#!/bin/bash
# round-2 GPU call 15: overlapped launches (programmatic dependent launch) on the sparse in-place storage; fast formatter
cd "$(dirname "$0")/.."
O=gpurun_out/r2c15; mkdir -p $O/w/out
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_outputs.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -5 $O/pytest_some.log
for ov in 0 1; do for cs in ldc pos bif; do for pr in f32 f64; do python tools/small_case.py --case $cs --precision $pr --overlap $ov --steps 400 --calls 2 | tail -1; done; done; done
python tools/small_case.py --case ldc --n 128 --precision f32 --overlap 0 --steps 200 --calls 2 | tail -1
python tools/small_case.py --case ldc --n 128 --precision f32 --overlap 1 --steps 200 --calls 2 | tail -1
( cd $O/w && LBM_TRACE=1 ../../../drivers/ldc > ldc.log 2> ldc.err; tail -2 ldc.log; cat ldc.err; LBM_TRACE=1 ../../../drivers/poiseuille > pos.log 2> pos.err; tail -1 pos.log; cat pos.err )
rm -rf $O/w
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
timeout 600 python tools/sparse_bench.py > $O/sparse_bench.json 2> $O/sparse_bench.err; tail -c 1500 $O/sparse_bench.json
