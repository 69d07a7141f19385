#!/bin/bash
# round-2, 2 GPUs, second call: e2e setup breakdown, parity of all transports after the arithmetic change
cd "$(dirname "$0")/.."
O=gpurun_out/r2m2b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in "LBM_AA=1" "LBM_SPARSE_AA=1" "LBM_P2P=0"; do
  env $mode timeout 300 $TR --master-port 29511 tools/mgpu_check.py > $O/mgpu_check_$(echo $mode | tr ' =' '__').log 2>&1
  echo "== $mode exit $?"; grep "\[mgpu\]" $O/mgpu_check_$(echo $mode | tr ' =' '__').log
done
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2.json 2> $O/bench_2.err
python -c "import json;d=json.loads(open('$O/bench_2.json').read().strip().split('\n')[-1]);print(d['value'],d['ms_per_step'],json.dumps(d['e2e']),json.dumps(d['parity_check']))"
tail -3 $O/bench_2.err
timeout 600 $TR --master-port 29513 tools/vessel_scale.py --size 512 --storage sparse_aa --steps 50 --verify > $O/vessel_2.json 2> $O/vessel_2.err; cat $O/vessel_2.json
timeout 600 $TR --master-port 29514 tools/vessel_scale.py --size 512 --storage sparse_aa --precision f32 --steps 50 > $O/vessel_2_f32.json 2>> $O/vessel_2.err; cat $O/vessel_2_f32.json
