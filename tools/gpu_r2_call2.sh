#!/bin/bash
# round-2 GPU call 2: GPU test suite with the group API, reference variants, bench with parity_check
cd "$(dirname "$0")/.."
O=gpurun_out/r2c2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench_f64.json 2> $O/bench_f64.err; tail -c 2500 $O/bench_f64.json; tail -5 $O/bench_f64.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 1200 $O/bench_ref.json
timeout 1800 python tools/capture_reference.py --variants gpurun_out/ref_variants > $O/variants.log 2>&1; tail -20 $O/variants.log
