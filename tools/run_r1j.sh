#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python tools/sparse_bench.py --steps 30 > gpurun_out/sparse_bench.json 2>&1; echo "sparse rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/sparse_bench.json'))
for k,v in d.items(): print(k, v if not isinstance(v,dict) else (round(v['mlups']), round(v['ms_per_step'],3), round(v['algorithmic_GBps']), round(v['device_GB'],1), round(v['fill'],3)))
"
python tools/sparse_bench.py --steps 30 --precision f32 --only sparse_ab 2>&1 | grep -E "mlups|algorithmic"
