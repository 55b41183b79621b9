#!/bin/bash
# what limits the strong-scaling leg at 8 GPUs: the per-rank share (125 k votes of a 10,000-key committee) on ONE GPU, kernel by kernel
set -u
mkdir -p gpurun_out
python bench.py --workload qc --committee 10000 --qcs 19 --votes-per-qc 6667 --steps 20 --warmup 3 > gpurun_out/r2_qc_125k.json 2>/dev/null; python tools/variants/lastjson.py gpurun_out/r2_qc_125k.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_qc_125k.csv python bench.py --workload qc --committee 10000 --qcs 19 --votes-per-qc 6667 --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_launches_qc_125k.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>vi: d[(r[ki].split('(')[0][-30:], r[gi])].append(float(r[vi].replace(',','')))
for k,v in d.items(): print("%-32s grid %-14s n=%3d median %.1f us"%(k[0],k[1],len(v),sorted(v)[len(v)//2]/1e3))
PY
