#!/bin/bash
# round-2 GPU call 28: fp32 sparse in-place kernels at 11 / 9 CTAs per SM
cd "$(dirname "$0")/.."
O=gpurun_out/r2c28; mkdir -p $O
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -3 $O/pytest_some.log
for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision f32 --steps 400 --calls 2 | tail -1; done
timeout 600 python tools/sparse_bench.py > $O/sparse_bench_f64.json 2> $O/sparse_bench.err
timeout 600 python tools/sparse_bench.py --precision f32 > $O/sparse_bench_f32.json 2>> $O/sparse_bench.err
python -c "
import json
for f in ('sparse_bench_f64','sparse_bench_f32'):
    d=json.load(open('$O/%s.json'%f)); print(f, {k:(round(v['mlups']),round(v['ms_per_step'],3),round(v['frac_of_measured_peak'],3)) for k,v in d.items() if isinstance(v,dict)}, d.get('same_fields'))
"
timeout 600 python bench.py --workload vessel --precision f32 --steps 50 --no-cpu > $O/bench_vessel_f32.json 2> $O/bench_vessel_f32.err; python -c "import json;d=json.loads(open('$O/bench_vessel_f32.json').read().strip().splitlines()[-1]);print(round(d['value']),d['roofline']['frac'],d['parity_check']['ok'],d['e2e']['value'])"
