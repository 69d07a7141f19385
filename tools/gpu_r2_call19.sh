#!/bin/bash
# round-2 GPU call 19: sparse in-place even step pulls before the lane test
cd "$(dirname "$0")/.."
O=gpurun_out/r2c19; mkdir -p $O
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_edge_cases_gpu.py tests/test_reference_outputs.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -4 $O/pytest_some.log
for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision f32 --steps 400 --calls 2 | tail -1; done
timeout 600 python tools/sparse_bench.py --only sparse_aa > $O/sparse_bench_f64.json 2> $O/sparse_bench.err; python -c "import json;d=json.load(open('$O/sparse_bench_f64.json'));print({k:(v['mlups'],v['ms_per_step'],v['frac_of_measured_peak']) for k,v in d.items() if isinstance(v,dict)})"
timeout 600 python tools/sparse_bench.py --only sparse_aa --precision f32 > $O/sparse_bench_f32.json 2>> $O/sparse_bench.err; python -c "import json;d=json.load(open('$O/sparse_bench_f32.json'));print({k:(v['mlups'],v['ms_per_step'],v['frac_of_measured_peak']) for k,v in d.items() if isinstance(v,dict)})"
