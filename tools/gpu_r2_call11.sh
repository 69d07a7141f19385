#!/bin/bash
# round-2 GPU call 11: explicit-fma FAST arithmetic (bitwise equality across kernel variants), fp32 two-segment kernel
cd "$(dirname "$0")/.."
O=gpurun_out/r2c11; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -8 $O/pytest.log
run() { python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --no-parity "$@" | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('$LBM_TAG', ' '.join(sys.argv[1:]), round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))" "$@"; }
LBM_TAG=default run --precision f32
LBM_TAG=x2off LBM_F32X2=0 run --precision f32
LBM_TAG=x2_5 LBM_B200_LIB=$PWD/variants/liblbm_x2_5.so run --precision f32
LBM_TAG=x2_4 LBM_B200_LIB=$PWD/variants/liblbm_x2_4.so run --precision f32
LBM_TAG=default run --precision f32 --storage ab
LBM_TAG=x2off LBM_F32X2=0 run --precision f32 --storage ab
LBM_TAG=default run --precision f64
LBM_TAG=default run --precision f64 --storage ab
for prec in f64 f32; do python tools/sparse_bench.py --n 512 --steps 50 --precision $prec --only sparse_aa | python -c "import json,sys;d=json.load(sys.stdin)['sparse_aa'];print('sparse_aa $prec', round(d['mlups']), round(d['ms_per_step'],3), round(d['frac_of_measured_peak'],4))"; done
timeout 900 python tools/compare_reference_runs.py > $O/reference_vs_ours_64.txt 2>&1; cat $O/reference_vs_ours_64.txt
