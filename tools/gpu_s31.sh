#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s31_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s31_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
rm -rf gpurun_out/drv_readme
timeout 600 python -m chain_b200.driver -s haagerupq -b p -l 6 -d 1600 -c 1e-8 -p 1e-2 -j 0,-1,0 -u 2 -q 0 -n 2 -m 5 --out gpurun_out/drv_readme --seed 1 2>&1 | tail -8
timeout 600 python -m chain_b200.driver -s golden -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 3 -m 5 --out gpurun_out/drv_cfg1 --seed 1 2>&1 | tail -6
