#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s23_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s23_pytest.log
for w in tiny cfg3 cfg2; do echo "$w: $(timeout 600 python tools/bond_update_times.py $w 2>&1 | tail -1 | cut -c1-220)"; done
timeout 600 python tools/sweep_bench.py golden p 24 800 -1 3 > gpurun_out/s23_sweep_golden.json 2> gpurun_out/s23_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s23_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
