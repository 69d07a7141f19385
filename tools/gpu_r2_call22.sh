#!/bin/bash
# round-2 GPU call 22: stopping rule of run_converge on the device (sparse in-place storage)
cd "$(dirname "$0")/.."
O=gpurun_out/r2c22; mkdir -p $O/w/out
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_outputs.py tests/test_drivers_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -8 $O/pytest_some.log
cd $O/w
for args in "" "--save 100000" "--f64"; do echo "== ldc $args"; LBM_TRACE=1 ../../../drivers/ldc $args > ldc.log 2> ldc.err; grep TOTAL ldc.log; cat ldc.err; done
for args in ""; do echo "== poiseuille $args"; LBM_TRACE=1 ../../../drivers/poiseuille $args > pos.log 2> pos.err; grep TOTAL pos.log; cat pos.err; done
cd ../../..; rm -rf $O/w
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
