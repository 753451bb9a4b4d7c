#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s34_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s34_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/s34_bench.json 2> gpurun_out/s34_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s34_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['dense_layout']['value'], d['roofline']['frac'], d['bond_update']['total_s'])"
tail -3 gpurun_out/s34_bench.err
