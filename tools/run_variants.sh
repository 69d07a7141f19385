export LBM_SPECULATIVE=2
for cfg in 4 5 6 7; do pr=f32
LBM_STEP_CFG=$cfg python bench.py --precision $pr --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg $cfg $pr', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))"
done
for cfg in 4 5; do pr=f64
LBM_STEP_CFG=$cfg python bench.py --precision $pr --steps 50 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg $cfg $pr', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))"
done
