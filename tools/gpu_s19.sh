#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bond_update_times.py cfg4s 2>&1 | tail -1 | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/s19_launches_update_cfg4s.csv python tools/bond_update_times.py cfg4s > gpurun_out/s19_ncu.log 2>&1
tail -1 gpurun_out/s19_ncu.log | cut -c1-300
