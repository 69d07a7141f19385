#!/bin/bash
# 2-GPU call: in-place storage across slabs (parity + bench), then one ncu --set full of the AA kernel
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
LBM_AA=1 timeout 300 $TR tools/mgpu_check.py 2>&1 | grep mgpu
LBM_P2P=1 LBM_SPARSE=1 timeout 300 $TR tools/mgpu_check.py 2>&1 | grep mgpu
timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu > gpurun_out/r01_scale3_2.json 2> gpurun_out/scale3_2.err; echo "bench2 rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-e2e --dims 1024 1024 1024 > gpurun_out/r01_scale3_1024_2.json 2> gpurun_out/scale3_1024_2.err; echo "bench1024 rc=$?"
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_short.json 2>&1 && \
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:k_step_dense -s 4 -c 2 -o gpurun_out/r01_aa_f64 -f \
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
for f in ('r01_scale3_2','r01_scale3_1024_2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['config']['storage'], d['config']['halo_exchange'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
