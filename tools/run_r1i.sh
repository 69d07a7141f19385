#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python tools/small_grid_probe.py > gpurun_out/small_grid_probe.txt 2>&1; cat gpurun_out/small_grid_probe.txt
python tools/sparse_bench.py --steps 30 > gpurun_out/sparse_bench.json 2>&1; echo "sparse rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/sparse_bench.json'))
for k,v in d.items(): print(k, v if not isinstance(v,dict) else (round(v['mlups']), round(v['ms_per_step'],3), round(v['algorithmic_GBps']), round(v['device_GB'],1), round(v['fill'],3)))
"
python tools/sparse_bench.py --steps 4 --only sparse_ab > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_sparse -s 6 -c 1 -o gpurun_out/r01_sparse_f64 -f python tools/sparse_bench.py --steps 4 --only sparse_ab > gpurun_out/ncu_sparse.log 2>&1
echo "ncu rc=$?"
