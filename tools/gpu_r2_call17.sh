#!/bin/bash
# round-2 GPU call 17: speculative path fetches node / wall words for mixed segments only
cd "$(dirname "$0")/.."
O=gpurun_out/r2c17; mkdir -p $O
timeout 900 python -m pytest tests/test_aa_gpu.py tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_mailbox_gpu.py tests/test_slab_gpu.py tests/test_sparse_gpu.py -m gpu -q -p no:cacheprovider -x > $O/pytest_some.log 2>&1; tail -5 $O/pytest_some.log
for pr in f64 f32; do for st in aa ab; do
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --precision $pr --storage $st > $O/bench_${pr}_$st.json 2> $O/bench_${pr}_$st.err; python -c "import json;d=json.loads(open('$O/bench_${pr}_$st.json').read().strip().split('\n')[-1]);print('$pr $st', d['value'],d['ms_per_step'],d['roofline']['frac'],d.get('parity_check',{}).get('max_rel_err'))"
done; done
