#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s33_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s33_pytest.log
for c in cfg1 cfg3 cfg2; do timeout 600 python tools/driver_times.py $c 2>&1 | tail -1 | cut -c1-420; done | tee gpurun_out/s33_driver_times.log
timeout 600 python tools/sweep_bench.py golden p 24 800 -1 3 > gpurun_out/s33_sweep_golden.json 2> gpurun_out/s33_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s33_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/s33_bench.json 2> gpurun_out/s33_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s33_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['bond_update']['total_s'], d['bond_update']['truncation_s'])"
