#!/bin/bash
mkdir -p gpurun_out
for cfg in "400 200 r" "1600 800 r" "3200 533 c"; do
  set -- $cfg
  for cps in 1 2; do
    CB200_TRI_CTAS_PER_SM=$cps ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s7_heig_$1_$cps.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s7_heig_$1_$cps.log 2>&1
    tail -1 gpurun_out/s7_heig_$1_$cps.log
  done
done
