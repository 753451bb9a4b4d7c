#!/bin/bash
mkdir -p gpurun_out
for cfg in "1600 800 r" "3200 533 c"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s29_heig_$1.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s29_heig_$1.log 2>&1
  python - gpurun_out/s29_heig_$1.csv <<'PY'
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for r in rows[hi+1:]:
    if len(r)>mv:
        try: v=float(r[mv].replace(',',''))
        except: continue
        tot+=v; k=r[kn].split('(')[0][:44]; agg[k][0]+=1; agg[k][1]+=v
print(sys.argv[1], "total %.2f ms"%(tot/1e6), {k:(v[0], round(v[1]/1e6,2)) for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:6]})
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s29_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s29_pytest.log
for w in cfg3 cfg2; do echo "$w: $(timeout 600 python tools/bond_update_times.py $w 2>&1 | tail -1 | cut -c1-200)"; done
