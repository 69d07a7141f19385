#!/bin/bash
# round-2 GPU call 9: two-level grid barrier in the persistent kernel; drivers vs reference programs
cd "$(dirname "$0")/.."
O=gpurun_out/r2c9; mkdir -p $O
timeout 900 python -m pytest tests/test_sparse_aa_gpu.py tests/test_reference_variants.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for prec in f32 f64; do for p in 0 1; do for cs in ldc pos bif; do python tools/small_case.py --case $cs --precision $prec --persistent $p --steps 400 --calls 2 | tail -1; done; done; done 2>&1 | tee $O/small_case.txt
timeout 600 python tools/small_grid_probe.py > $O/small_grid_probe.txt 2>&1; cat $O/small_grid_probe.txt
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
timeout 600 python bench.py --steps 50 --warmup 5 --storage sparse_aa --no-cpu > $O/bench_f64_sparse_aa.json 2> $O/err1.txt; python -c "import json;d=json.loads(open('$O/bench_f64_sparse_aa.json').read().strip().split('\n')[-1]);print('LDC512 f64 sparse_aa', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_check'])"
timeout 600 python bench.py --steps 50 --warmup 5 --storage sparse_aa --precision f32 --no-cpu > $O/bench_f32_sparse_aa.json 2> $O/err2.txt; python -c "import json;d=json.loads(open('$O/bench_f32_sparse_aa.json').read().strip().split('\n')[-1]);print('LDC512 f32 sparse_aa', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_check'])"
timeout 600 python bench.py --steps 50 --warmup 5 --precision f32 --no-cpu > $O/bench_f32_aa.json 2> $O/err3.txt; python -c "import json;d=json.loads(open('$O/bench_f32_aa.json').read().strip().split('\n')[-1]);print('LDC512 f32 aa', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_check'])"
tail -3 $O/err1.txt $O/err2.txt $O/err3.txt
