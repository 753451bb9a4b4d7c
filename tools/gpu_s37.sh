#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python tools/bond_update_times.py cfg2 > gpurun_out/s37_bond_update.json 2> gpurun_out/s37_bond_update.err; echo "plain rc=$?"; cat gpurun_out/s37_bond_update.json | cut -c1-300
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s37_launches_bond_update_cfg2.csv python tools/bond_update_times.py cfg2 > gpurun_out/s37_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/s37_ncu.log | cut -c1-300
wc -l gpurun_out/s37_launches_bond_update_cfg2.csv
