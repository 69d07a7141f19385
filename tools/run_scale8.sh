#!/bin/bash
# one 8-GPU call: parity at 8 ranks for the three storages, then the scaling lines
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
LBM_AA=1 timeout 300 $TR --nproc-per-node 8 --master-port 29501 tools/mgpu_check.py 2>&1 | grep mgpu
LBM_P2P=1 LBM_SPARSE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29502 tools/mgpu_check.py 2>&1 | grep mgpu
timeout 400 $TR --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/r01_scale3_8.json
timeout 400 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/r01_scale3_4.json
timeout 400 $TR --nproc-per-node 8 --master-port 29505 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --dims 1024 1024 1024 2>&1 | tail -1 > gpurun_out/r01_scale3_1024_8.json
timeout 400 $TR --nproc-per-node 8 --master-port 29506 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --precision f32 --dims 1024 1024 1024 2>&1 | tail -1 > gpurun_out/r01_scale3_1024_8_f32.json
timeout 500 $TR --nproc-per-node 8 --master-port 29507 tools/vessel_scale.py --size 1024 --k 4 --steps 30 2>&1 | tail -1 > gpurun_out/r01_vessel3_1024_8.json
for f in gpurun_out/r01_scale3_8.json gpurun_out/r01_scale3_4.json gpurun_out/r01_scale3_1024_8.json gpurun_out/r01_scale3_1024_8_f32.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['config']['storage'], d['config'].get('halo_exchange'), d['gpu_launches'], round(d['roofline']['frac'],4))
except Exception as e: print(sys.argv[1], 'ERR', e, open(sys.argv[1]).read()[-300:])
PY
done
cat gpurun_out/r01_vessel3_1024_8.json
