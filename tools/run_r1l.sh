#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/small_grid_probe.py 2>&1 | tee gpurun_out/small_grid_probe.txt
echo "--- LBM_SPECULATIVE=1"
LBM_SPECULATIVE=1 python tools/small_grid_probe.py 2>&1 | tee gpurun_out/small_grid_probe_spec.txt
python tools/compare_reference_runs.py > gpurun_out/r01_reference_vs_ours_64.txt 2>&1; cat gpurun_out/r01_reference_vs_ours_64.txt | head -12
