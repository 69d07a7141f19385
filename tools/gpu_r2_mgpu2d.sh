#!/bin/bash
# round-2, 2 GPUs: in-place dense storage over NCCL send/recv (staged mailboxes, no peer mapping): parity and speed
cd "$(dirname "$0")/.."
O=gpurun_out/r2m2d; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
LBM_AA=1 LBM_STAGED=1 timeout 300 $TR --master-port 29511 tools/mgpu_check.py > $O/mgpu_check_staged.log 2>&1; echo "exit $?"; grep "\[mgpu\]" $O/mgpu_check_staged.log; tail -3 $O/mgpu_check_staged.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --halo nccl --steps 50 --no-cpu > $O/bench_2_nccl_aa.json 2> $O/bench_2_nccl_aa.err; echo "rc=$?"
python -c "import json;d=json.loads(open('$O/bench_2_nccl_aa.json').read().strip().split('\n')[-1]);print(d['value'],d['ms_per_step'],d['config']['halo_exchange'],d['config']['storage'],d['e2e']['value'],json.dumps(d['parity_check']['cases']), d['gpu_launches'])"
tail -3 $O/bench_2_nccl_aa.err
