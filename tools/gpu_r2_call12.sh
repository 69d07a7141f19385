#!/bin/bash
# round-2 GPU call 12: where the drivers' time goes; persistent kernel under ncu (source level)
cd "$(dirname "$0")/.."
O=gpurun_out/r2c12; mkdir -p $O/w/out
( cd $O/w && for i in 1 2; do LBM_TRACE=1 ../../../drivers/ldc > ldc.log 2> ldc.err; tail -2 ldc.log; cat ldc.err; done; LBM_TRACE=1 ../../../drivers/poiseuille > pos.log 2> pos.err; tail -1 pos.log; cat pos.err; LBM_TRACE=1 ../../../drivers/ldc --storage aa > ldcaa.log 2> ldcaa.err; tail -1 ldcaa.log; cat ldcaa.err )
rm -rf $O/w
for p in 0 1; do python tools/small_case.py --case ldc --precision f32 --persistent $p --steps 400 --calls 2 | tail -1; done
CMD="python tools/small_case.py --case ldc --persistent 1 --steps 40 --calls 1"
$CMD > $O/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sparse_aa_persist -s 1 -c 1 -o $O/persist2 $CMD > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
