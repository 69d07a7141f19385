#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
LBM_P2P=1 LBM_SPARSE=1 timeout 300 $TR --nproc-per-node 4 --master-port 29502 tools/mgpu_check.py 2>&1 | grep "mgpu\|Error\|unavailable" | head
LBM_AA=1 timeout 300 $TR --nproc-per-node 4 --master-port 29503 tools/mgpu_check.py 2>&1 | grep "mgpu\|Error\|unavailable" | head
timeout 400 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | cut -c1-300
python -m pytest tests -m gpu -x -q -k "multigpu or slab" 2>&1 | tail -2
