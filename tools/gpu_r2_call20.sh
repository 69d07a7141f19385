#!/bin/bash
# round-2 GPU call 20: bench.py --workload vessel (BASELINE config 5) on one GPU, fp64 and fp32; geo.txt reader through the drivers' tests
cd "$(dirname "$0")/.."
O=gpurun_out/r2c20; mkdir -p $O
timeout 900 python bench.py --workload vessel --steps 50 > $O/bench_vessel_f64.json 2> $O/bench_vessel_f64.err; echo "rc=$?"; tail -3 $O/bench_vessel_f64.err
timeout 900 python bench.py --workload vessel --steps 50 --precision f32 --no-cpu > $O/bench_vessel_f32.json 2> $O/bench_vessel_f32.err; echo "rc=$?"
python -c "
import json
for f in ('bench_vessel_f64','bench_vessel_f32'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['config']['storage'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['roofline']['traffic'], 'e2e', round(d['e2e']['value']), d['e2e']['phases'], d['e2e']['h2d_bytes_per_step'], d['parity_check'], d.get('gpu_launches'))
"
timeout 900 python -m pytest tests/test_reference_outputs.py tests/test_drivers_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
