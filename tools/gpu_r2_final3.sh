#!/bin/bash
# round-2: full GPU suite, smoke, default bench on the final library
cd "$(dirname "$0")/.."
O=gpurun_out/r2final3; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench_f64.json 2> $O/bench_f64.err; echo "bench rc=$?"
timeout 900 python bench.py --precision f32 --no-cpu > $O/bench_f32.json 2> $O/bench_f32.err
python -c "
import json
for f in ('bench_f64','bench_f32'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['config']['storage'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e']['phases'], d.get('gpu_launches'), d['parity_check']['ok'])
"
