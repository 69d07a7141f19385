#!/bin/bash
# round-2 GPU call 21: where do the drivers' loops spend their time (dumps on / off), vessel e2e with small staging groups
cd "$(dirname "$0")/.."
O=gpurun_out/r2c21; mkdir -p $O/w/out
cd $O/w
for args in "" "--save 100000" "--f64" "--f64 --save 100000"; do echo "== ldc $args"; LBM_TRACE=1 ../../../drivers/ldc $args > ldc.log 2> ldc.err; tail -2 ldc.log | head -1; cat ldc.err; done
for args in "" "--save 100000"; do echo "== poiseuille $args"; LBM_TRACE=1 ../../../drivers/poiseuille $args > pos.log 2> pos.err; grep TOTAL pos.log; cat pos.err; done
cp ../../../tests/golden/geo_bif.txt geo.txt 2>/dev/null; ls ../../../tests/golden | head -20
cd ../../..; rm -rf $O/w
timeout 600 python bench.py --workload vessel --steps 50 --no-cpu --no-parity | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value']),d['e2e']['value'],d['e2e']['phases'])"
