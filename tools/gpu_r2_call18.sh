#!/bin/bash
# round-2 GPU call 18: ncu --set full of the fp32 in-place step, product kernel against the boundary-free kbench pair
cd "$(dirname "$0")/.."
O=gpurun_out/r2c18; mkdir -p $O
tools/kbench 512 0 64 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aa_ -s 6 -c 2 -o $O/kbench_aa_f32 -f tools/kbench 512 0 64 1 > $O/ncu_kbench.log 2>&1; tail -3 $O/ncu_kbench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_dense -s 4 -c 2 -o $O/dense_aa_f32 -f python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --no-parity --precision f32 > $O/ncu_dense.log 2>&1; tail -3 $O/ncu_dense.log
ls -la $O
