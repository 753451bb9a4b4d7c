#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches_update_cfg2.csv python tools/bond_update_times.py cfg2 > gpurun_out/s5_ncu_update.log 2>&1
tail -2 gpurun_out/s5_ncu_update.log | cut -c1-300
