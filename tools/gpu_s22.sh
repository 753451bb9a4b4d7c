#!/bin/bash
mkdir -p gpurun_out
for cfg in "400 200 r" "1600 800 r" "3200 533 c"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s22_heig_$1.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s22_heig_$1.log 2>&1
  python - gpurun_out/s22_heig_$1.csv <<'PY'
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
t=0.0; tot=0.0
for r in rows[hi+1:]:
    if len(r)>mv:
        try: v=float(r[mv].replace(',',''))
        except: continue
        tot+=v
        if 'tridiag_panel' in r[kn]: t+=v
print(sys.argv[1], "total %.2f ms tridiag %.2f ms"%(tot/1e6, t/1e6))
PY
done
