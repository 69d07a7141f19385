#!/bin/bash
# round-2 GPU call 26: ncu --set full of the final sparse in-place kernels (vessel bundle 512^3), fp64 and fp32
cd "$(dirname "$0")/.."
O=gpurun_out/r2c26; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa_ -s 10 -c 2 -o $O/sparse_aa_f64 -f python tools/sparse_bench.py --only sparse_aa --steps 12 > $O/ncu_f64.log 2>&1; tail -2 $O/ncu_f64.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa_ -s 10 -c 2 -o $O/sparse_aa_f32 -f python tools/sparse_bench.py --only sparse_aa --steps 12 --precision f32 > $O/ncu_f32.log 2>&1; tail -2 $O/ncu_f32.log
ls -la $O
