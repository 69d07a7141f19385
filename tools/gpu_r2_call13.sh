#!/bin/bash
# round-2 GPU call 13: mailbox mode on virtual slabs; full suite; drivers trace
cd "$(dirname "$0")/.."
O=gpurun_out/r2c13; mkdir -p $O/w/out
timeout 600 python -m pytest tests/test_mailbox_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_mail.log 2>&1; tail -15 $O/pytest_mail.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
( cd $O/w && LBM_TRACE=1 ../../../drivers/ldc > ldc.log 2> ldc.err; tail -2 ldc.log; cat ldc.err; LBM_TRACE=1 ../../../drivers/poiseuille > pos.log 2> pos.err; tail -1 pos.log; cat pos.err )
rm -rf $O/w
for p in 0 1; do for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision f32 --persistent $p --steps 400 --calls 2 | tail -1; done; done
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
