#!/bin/bash
# ncu launch list + --set full summaries of the block-Jacobi SVD kernels on a pathfinder-size block
OUT=gpurun_out
mkdir -p $OUT
B="python tools/bench_svd.py --only 9 ${SVD_CASE:---big4}"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file $OUT/r01_svd_launches.csv $B > $OUT/ncu_svd_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bj_ --launch-skip 300 -c 4 -f -o /tmp/svd $B > $OUT/ncu_svd.log 2>&1
python profiles/ncu_summary.py /tmp/svd.ncu-rep > $OUT/r01_ncu_svd.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:hh_.*apply --launch-skip 400 -c 2 -f -o /tmp/svdhh $B > $OUT/ncu_svdhh.log 2>&1
python profiles/ncu_summary.py /tmp/svdhh.ncu-rep >> $OUT/r01_ncu_svd.txt 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r01_svd_launches.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i; break
ix={h:i for i,h in enumerate(hdr)}
from collections import defaultdict
agg=defaultdict(lambda:[0,0.0])
for r in rows[start+1:]:
    if len(r)<len(hdr): continue
    k=r[ix['Kernel Name']].split('(')[0][:50]
    agg[k][0]+=1; agg[k][1]+=float(r[ix['Metric Value']].replace(',',''))
for k,v in agg.items(): print(k, v[0], 'launches', v[1]/1e3, 'us total', v[1]/v[0]/1e3, 'us avg')
PY
