#!/bin/bash
# round-2 GPU call 6: persistent small-grid path, drivers vs the reference programs, full GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out/r2c6; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 600 python tools/small_grid_probe.py > $O/small_grid_probe.txt 2>&1; cat $O/small_grid_probe.txt
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
timeout 900 python tools/selfcheck.py > $O/selfcheck.log 2>&1; cat $O/selfcheck.log
