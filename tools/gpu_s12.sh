#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s12_pytest.log
for cfg in "400 200 r" "1600 800 r" "3200 533 c"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s12_heig_$1.csv python tools/heig_bench.py $1 $2 $3 > gpurun_out/s12_heig_$1.log 2>&1
  tail -1 gpurun_out/s12_heig_$1.log
done
timeout 900 python tools/sweep_bench.py golden p 24 800 -1 2 > gpurun_out/s12_sweep_golden.json 2> gpurun_out/s12_sweep_golden.err; cut -c1-1500 gpurun_out/s12_sweep_golden.json
timeout 1500 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 3 > gpurun_out/s12_sweep_hq.json 2> gpurun_out/s12_sweep_hq.err; cut -c1-2000 gpurun_out/s12_sweep_hq.json; tail -3 gpurun_out/s12_sweep_hq.err
