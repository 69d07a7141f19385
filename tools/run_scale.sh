TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
LBM_P2P=1 LBM_SPARSE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29501 tools/mgpu_check.py 2>&1 | grep mgpu
LBM_P2P=1 timeout 600 $TR --nproc-per-node 8 --master-port 29502 tools/mgpu_check.py 2>&1 | grep "ALL\|FAILED"
for h in p2p nccl; do
timeout 600 $TR --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --halo $h --dims 1024 1024 1024 2>&1 | tail -1 > gpurun_out/scale2_1024_8_$h.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29504 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --halo p2p 2>&1 | tail -1 > gpurun_out/scale2_weak_8_p2p.json
timeout 600 $TR --nproc-per-node 4 --master-port 29505 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu --no-e2e --halo p2p 2>&1 | tail -1 > gpurun_out/scale2_weak_4_p2p.json
timeout 600 $TR --nproc-per-node 2 --master-port 29506 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-e2e --halo p2p 2>&1 | tail -1 > gpurun_out/scale2_weak_2_p2p.json
python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/scale2_weak_1.json
timeout 600 $TR --nproc-per-node 8 --master-port 29507 tools/vessel_scale.py --size 1024 --k 4 --steps 30 2>&1 | tail -1 > gpurun_out/vessel_1024_8.json
for f in gpurun_out/scale2_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1], d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['config']['workload'][:46], d['config'].get('halo_exchange'), d['gpu_launches'])
PY
done
cat gpurun_out/vessel_1024_8.json
