#!/bin/bash
mkdir -p gpurun_out
for cap in 0 8 16 64; do
  for cfg in "400 200 r" "1600 800 r"; do
    set -- $cfg
    CB200_TRI_MAX_CTAS=$cap timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s21_heig_$1_$cap.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s21_heig_$1_$cap.log 2>&1
    python - gpurun_out/s21_heig_$1_$cap.csv $cap <<'PY'
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
t=0.0
for r in rows[hi+1:]:
    if len(r)>mv and 'tridiag_panel' in r[kn]:
        try: t+=float(r[mv].replace(',',''))
        except: pass
print(sys.argv[1], "cap", sys.argv[2], "tridiag %.2f ms"%(t/1e6))
PY
  done
done
timeout 600 python tools/sweep_bench.py golden p 24 800 -1 3 > gpurun_out/s21_sweep_golden.json 2> gpurun_out/s21_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s21_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
