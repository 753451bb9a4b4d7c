#!/bin/bash
mkdir -p gpurun_out
for fs in 1 4 8 16; do
  CB200_TRI_FSPLIT=$fs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s30_heig_3200_$fs.csv python tools/heig_bench.py 3200 533 c > gpurun_out/s30_heig_3200_$fs.log 2>&1
  python - gpurun_out/s30_heig_3200_$fs.csv $fs <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
t=0
for r in rows[hi+1:]:
    if len(r)>mv and 'tridiag_panel' in r[kn]:
        try: t+=float(r[mv].replace(',',''))
        except: pass
print("fsplit", sys.argv[2], "tridiag n=3200c %.2f ms"%(t/1e6), open(sys.argv[1].replace('.csv','.log')).read().strip().splitlines()[-1][-60:])
PY
done
