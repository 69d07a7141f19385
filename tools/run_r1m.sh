#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python tools/small_grid_kernel_time.py && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_launches.csv python tools/small_grid_kernel_time.py > gpurun_out/ncu_small.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/small_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[i0]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); g=hdr.index('Grid Size')
d=collections.defaultdict(list)
for r in rows[i0+1:]:
    if len(r)>mv and 'k_step' in r[kn]: d[(r[kn][:60], r[g])].append(float(r[mv].replace(',','')))
for k,v in d.items(): print(k, len(v), 'median ns', sorted(v)[len(v)//2])
PY
