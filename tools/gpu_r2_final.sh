#!/bin/bash
# round-2 final evidence on one GPU: full GPU suite, smoke, default / fp32 bench, ncu launch list of the default bench,
# ncu --set full of the fp64 in-place kernel, sparse bench, self-checking build
cd "$(dirname "$0")/.."
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_f64.json 2> $O/bench_f64.err; echo "bench rc=$?"
timeout 900 python bench.py --precision f32 --no-cpu > $O/bench_f32.json 2> $O/bench_f32.err
python -c "
import json
for f in ('bench_f64','bench_f32'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['config']['storage'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e']['phases'], d.get('cpu_baseline'), d.get('gpu_launches'), d.get('clocks'))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_f64_512_aa.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_dense -s 4 -c 2 -o $O/dense_aa_f64 -f python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-parity > $O/ncu_dense_f64.log 2>&1; echo "ncu full rc=$?"
timeout 600 python tools/sparse_bench.py > $O/sparse_bench_f64.json 2> $O/sparse_bench.err
timeout 600 python tools/sparse_bench.py --precision f32 > $O/sparse_bench_f32.json 2>> $O/sparse_bench.err
python -c "
import json
for f in ('sparse_bench_f64','sparse_bench_f32'):
    d=json.load(open('$O/%s.json'%f)); print(f, {k:(round(v['mlups']),round(v['ms_per_step'],3),round(v['frac_of_measured_peak'],3)) for k,v in d.items() if isinstance(v,dict)}, d.get('same_fields'))
"
timeout 900 python tools/selfcheck.py > $O/selfcheck.log 2>&1; echo "selfcheck rc=$?"; tail -3 $O/selfcheck.log; cp gpurun_out/selfcheck/*.log $O/ 2>/dev/null
