#!/bin/bash
# round-2, 8 GPUs (one call, charged 8x: every step under a short timeout): the bench at 8 and 4 ranks, the 1024^3
# vessel bundle on the sparse in-place storage, parity of both in-place storages at 8 ranks, the C multi-GPU driver
cd "$(dirname "$0")/.."
O=gpurun_out/r2m8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().split('\n')[-1]);print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','dtype')}, json.dumps(d.get('e2e',{}).get('value')), json.dumps(d.get('e2e',{}).get('phases')), (d.get('parity_check') or {}).get('ok'))" 2>&1 | tail -2; }
timeout 240 $TR8 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > $O/bench_8.json 2> $O/bench_8.err; echo "bench8 exit $?"; show $O/bench_8.json; tail -2 $O/bench_8.err
timeout 240 $TR4 --master-port 29522 bench.py --gpus 4 --steps 50 --warmup 5 > $O/bench_4.json 2> $O/bench_4.err; echo "bench4 exit $?"; show $O/bench_4.json
timeout 240 $TR8 --master-port 29523 tools/vessel_scale.py --size 1024 --storage sparse_aa --steps 50 > $O/vessel_1024_8.json 2> $O/vessel_8.err; echo "vessel exit $?"; cat $O/vessel_1024_8.json; tail -2 $O/vessel_8.err
timeout 240 $TR8 --master-port 29524 tools/vessel_scale.py --size 1024 --storage sparse_aa --precision f32 --steps 50 > $O/vessel_1024_8_f32.json 2>> $O/vessel_8.err; cat $O/vessel_1024_8_f32.json
for mode in "LBM_AA=1" "LBM_SPARSE_AA=1"; do
  env $mode timeout 120 $TR8 --master-port 29525 tools/mgpu_check.py > $O/mgpu_check_$(echo $mode | tr ' =' '__').log 2>&1
  echo "== $mode exit $?"; grep "\[mgpu\]" $O/mgpu_check_$(echo $mode | tr ' =' '__').log
done
timeout 240 $TR8 --master-port 29526 tools/vessel_scale.py --size 512 --storage sparse_aa --steps 30 --verify > $O/vessel_512_8.json 2>> $O/vessel_8.err; cat $O/vessel_512_8.json
timeout 240 $TR8 --master-port 29527 bench.py --gpus 8 --steps 50 --warmup 5 --precision f32 > $O/bench_8_f32.json 2> $O/bench_8_f32.err; show $O/bench_8_f32.json
timeout 120 drivers/ldc_mgpu --slabs 8 --n 256 --steps 1000 --save 1000 --out $O > $O/ldc_mgpu.log 2>&1; tail -3 $O/ldc_mgpu.log; rm -f $O/lid_*.vtk $O/CONVERGENCE.log
