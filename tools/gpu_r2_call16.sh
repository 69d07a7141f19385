#!/bin/bash
# round-2 GPU call 16: overlapped launches on the dense in-place storage too; fp32 vector-access variants (kbench)
cd "$(dirname "$0")/.."
O=gpurun_out/r2c16; mkdir -p $O
timeout 900 python -m pytest tests/test_aa_gpu.py tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_mailbox_gpu.py tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -5 $O/pytest_some.log
for ov in 0 1; do for pr in f32 f64; do python tools/small_case.py --case ldc --storage aa --precision $pr --overlap $ov --steps 400 --calls 2 | tail -1; done; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_f64.json 2> $O/bench_f64.err; python -c "import json;d=json.loads(open('$O/bench_f64.json').read().strip().split('\n')[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['parity_check'])"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --precision f32 > $O/bench_f32.json 2> $O/bench_f32.err; python -c "import json;d=json.loads(open('$O/bench_f32.json').read().strip().split('\n')[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'])"
timeout 300 tools/kbench 512 0 > $O/kbench_f32.txt 2>&1; grep -v "step b\|2 rows" $O/kbench_f32.txt
