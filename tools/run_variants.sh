python -m pytest tests/test_aa_gpu.py -x -q -m gpu 2>&1 | tail -8
for st in ab aa; do for pr in f64 f32; do
python bench.py --storage $st --precision $pr --steps 50 --warmup 6 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$st $pr', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))"
done; done
