#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sparse or slab or multigpu or edge or parity" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for spw in 1 2 4 8; do
  echo "SPW=$spw"
  LBM_SPW=$spw python tools/sparse_bench.py --steps 30 --only sparse_ab 2>&1 | grep -E '"mlups"|"ms_per_step"|algorithmic' | tr -d '\n'; echo
done
LBM_SPW=4 python tools/sparse_bench.py --steps 30 --precision f32 --only sparse_ab 2>&1 | grep -E '"mlups"|algorithmic' | tr -d '\n'; echo
LBM_SPW=1 python tools/sparse_bench.py --steps 30 --precision f32 --only sparse_ab 2>&1 | grep -E '"mlups"|algorithmic' | tr -d '\n'; echo
