#!/bin/bash
# round-2 final bench lines (median-of-3 end-to-end), reference arm
cd "$(dirname "$0")/.."
O=gpurun_out/r2final2; mkdir -p $O
timeout 900 python bench.py > $O/bench_f64.json 2> $O/bench_f64.err; echo "bench rc=$?"
timeout 900 python bench.py --precision f32 --no-cpu > $O/bench_f32.json 2> $O/bench_f32.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 900 $O/bench_ref.json
python -c "
import json
for f in ('bench_f64','bench_f32'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['config']['storage'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['roofline']['traffic'], 'e2e', round(d['e2e']['value']), d['e2e']['phases'], d.get('gpu_launches'))
"
