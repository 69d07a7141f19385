#!/bin/bash
# round-2 GPU call 7: persistent kernel with few large CTAs; ncu of the sparse in-place kernels (own numbering)
cd "$(dirname "$0")/.."
O=gpurun_out/r2c7; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 600 python tools/small_grid_probe.py > $O/small_grid_probe.txt 2>&1; cat $O/small_grid_probe.txt
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
for prec in f64 f32; do
CMD="python tools/sparse_bench.py --n 512 --steps 4 --precision $prec --only sparse_aa"
$CMD > $O/plain_$prec.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa -s 5 -c 2 -o $O/sparse_aa2_$prec $CMD > $O/ncu_$prec.log 2>&1
tail -2 $O/ncu_$prec.log
done
