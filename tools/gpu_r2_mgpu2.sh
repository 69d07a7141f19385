#!/bin/bash
# round-2, 2 GPUs: the multi-process path (flag handshake in peer memory, lbm_slab_step) -- parity, bench, scaling pieces
cd "$(dirname "$0")/.."
O=gpurun_out/r2m2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in "LBM_AA=1" "LBM_SPARSE_AA=1" "LBM_P2P=1" "LBM_P2P=1 LBM_SPARSE=1" "LBM_P2P=0"; do
  env $mode timeout 300 $TR --master-port 29511 tools/mgpu_check.py > $O/mgpu_check_$(echo $mode | tr ' =' '__').log 2>&1
  echo "== $mode exit $?"; grep "\[mgpu\]" $O/mgpu_check_$(echo $mode | tr ' =' '__').log
done
timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_2.json 2> $O/bench_2.err; tail -c 3000 $O/bench_2.json; tail -5 $O/bench_2.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > $O/bench_1.json 2> $O/bench_1.err; tail -c 600 $O/bench_1.json
timeout 600 $TR --master-port 29513 tools/vessel_scale.py --size 512 --storage sparse_aa --steps 50 --verify > $O/vessel_2.json 2> $O/vessel_2.err; cat $O/vessel_2.json; tail -3 $O/vessel_2.err
timeout 300 drivers/ldc_mgpu --slabs 2 --n 128 --steps 2000 --save 1000 --out $O > $O/ldc_mgpu.log 2>&1; tail -3 $O/ldc_mgpu.log; rm -f $O/lid_*.vtk
