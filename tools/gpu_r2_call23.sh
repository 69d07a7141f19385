#!/bin/bash
# round-2 GPU call 23: step kernels loaded in lbm_initialize (lazy module loading no longer inside the time loop)
cd "$(dirname "$0")/.."
O=gpurun_out/r2c23; mkdir -p $O/w/out
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_outputs.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -4 $O/pytest_some.log
python tools/small_case.py --case ldc --precision f32 --no-warm --converge 10001 --tol 1e-6 | tail -2
python tools/setup_probe.py --rounds 2 --storage sparse_aa --n 64 --precision f32
python tools/setup_probe.py --rounds 2
cd $O/w
for args in "" "--save 100000"; do echo "== ldc $args"; LBM_TRACE=1 ../../../drivers/ldc $args > ldc.log 2> ldc.err; grep TOTAL ldc.log; cat ldc.err; done
cd ../../..; rm -rf $O/w
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
