#!/bin/bash
# round-2 GPU call 14: one launch per step by default again, residual summed per CTA; drivers, small grids
cd "$(dirname "$0")/.."
O=gpurun_out/r2c14; mkdir -p $O/w/out
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_outputs.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -5 $O/pytest_some.log
( cd $O/w && LBM_TRACE=1 ../../../drivers/ldc > ldc.log 2> ldc.err; tail -2 ldc.log; cat ldc.err; LBM_TRACE=1 ../../../drivers/poiseuille > pos.log 2> pos.err; tail -1 pos.log; cat pos.err; LBM_TRACE=1 ../../../drivers/bifurcation > bif.log 2> bif.err; tail -1 bif.log; cat bif.err )
rm -rf $O/w
for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision f32 --steps 400 --calls 2 | tail -1; done
timeout 600 python tools/small_grid_probe.py > $O/small_grid.txt 2>&1; cat $O/small_grid.txt
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
