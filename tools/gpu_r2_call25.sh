#!/bin/bash
# round-2 GPU call 25: flag planes copied straight into the padded device rows (geo_pre); vessel set-up phases
cd "$(dirname "$0")/.."
O=gpurun_out/r2c25; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python tools/setup_probe.py --workload vessel --storage sparse_aa --rounds 2
timeout 600 python bench.py --workload vessel --steps 50 --no-cpu --no-parity | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value']),d['e2e']['value'],d['e2e']['phases'])"
