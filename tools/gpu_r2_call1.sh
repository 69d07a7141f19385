#!/bin/bash
# round-2 GPU call 1: full GPU test suite, sanitizer logs, reference variants, a bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt; free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log
timeout 1200 bash tools/run_sanitizer.sh > $O/sanitizer.log 2>&1
timeout 1500 python tools/capture_reference.py --variants gpurun_out/ref_variants > $O/variants.log 2>&1; tail -20 $O/variants.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench_f64.json 2> $O/bench_f64.err; tail -c 1500 $O/bench_f64.json
