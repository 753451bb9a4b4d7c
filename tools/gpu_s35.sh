#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s35_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s35_pytest.log
timeout 600 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 3 > gpurun_out/s35_sweep_hq.json 2> gpurun_out/s35_sweep_hq.err; echo "sweep rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s35_sweep_hq.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
timeout 300 python tools/sweep_bench.py golden p 24 800 -1 3 > gpurun_out/s35_sweep_golden.json 2> gpurun_out/s35_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s35_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
for c in cfg1 cfg2; do timeout 300 python tools/driver_times.py $c 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/s35_driver_times.log
