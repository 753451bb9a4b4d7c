#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q > gpurun_out/s2_pytest_shard.log 2>&1; echo "shard pytest rc=$?"; tail -15 gpurun_out/s2_pytest_shard.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shard.py > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s2_pytest.log
for w in cfg3 tiny cfg2; do echo "$w: $(python tools/bond_update_times.py $w 2>&1 | tail -1)"; done | tee gpurun_out/s2_update.log
