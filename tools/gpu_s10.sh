#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s10_pytest.log
for cfg in "400 200 r" "1600 800 r" "3200 533 c"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s10_heig_$1.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s10_heig_$1.log 2>&1
  tail -1 gpurun_out/s10_heig_$1.log
done
for c in cfg1 cfg3 cfg2; do timeout 900 python tools/driver_times.py $c 2>&1 | tail -1; done | tee gpurun_out/s10_driver_times.log
for w in tiny cfg3 cfg2; do echo "$w: $(timeout 600 python tools/bond_update_times.py $w 2>&1 | tail -1 | cut -c1-220)"; done | tee gpurun_out/s10_update.log
