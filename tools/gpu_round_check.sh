#!/bin/bash
# one gpurun call: GPU tests, default bench, then the ncu launch list of the same bench command
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --precision f32 --no-cpu > gpurun_out/bench_f32.json 2>> gpurun_out/bench_default.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_f64_512_aa.csv \
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu rc=$?"
python -c "
import json
for f in ('bench_default','bench_f32'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['config']['storage'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e']['phases'])
"
