set -x
python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
$TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_weak_$N.json
$TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-e2e --dims 1024 1024 1024 2>&1 | tail -1 > gpurun_out/scale_1024_$N.json
done
python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/scale_weak_1.json
for f in gpurun_out/scale_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read()); print('$f', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['config']['workload'][:44], d.get('e2e') and round(d['e2e']['value']))"; done
