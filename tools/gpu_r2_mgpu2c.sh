#!/bin/bash
# round-2, 2 GPUs, third call: mailbox transport across processes (dense in-place), e2e setup breakdown
cd "$(dirname "$0")/.."
O=gpurun_out/r2m2c; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in "LBM_AA=1" "LBM_SPARSE_AA=1"; do
  env $mode timeout 300 $TR --master-port 29511 tools/mgpu_check.py > $O/mgpu_check_$(echo $mode | tr ' =' '__').log 2>&1
  echo "== $mode exit $?"; grep "\[mgpu\]" $O/mgpu_check_$(echo $mode | tr ' =' '__').log
done
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2.json 2> $O/bench_2.err
python -c "import json;d=json.loads(open('$O/bench_2.json').read().strip().split('\n')[-1]);print(d['value'],d['ms_per_step'],json.dumps(d['e2e']),json.dumps(d['parity_check']),d.get('config'))"
tail -3 $O/bench_2.err
