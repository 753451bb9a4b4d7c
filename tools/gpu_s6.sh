#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "eigensolver" > gpurun_out/s6_pytest_heig.log 2>&1; echo "heig pytest rc=$?"; tail -15 gpurun_out/s6_pytest_heig.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s6_pytest.log
for mode in blocked unblocked; do for w in tiny cfg3 cfg2; do echo "$mode $w: $(CB200_TRIDIAG=$mode timeout 600 python tools/bond_update_times.py $w 2>&1 | tail -1 | cut -c1-220)"; done; done | tee gpurun_out/s6_update.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s6_launches_update_cfg2.csv python tools/bond_update_times.py cfg2 > gpurun_out/s6_ncu_update.log 2>&1
