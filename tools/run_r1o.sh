#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/small_grid_kernel_time.py
python bench.py --steps 50 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
python tools/sparse_bench.py --steps 30 --only sparse_ab 2>&1 | grep -E '"mlups"' 
