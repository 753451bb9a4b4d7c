#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/s38_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/s38_pytest.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/s38_pytest.log | cut -c1-300; CB200_WY_UPFRONT=0 timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k eigensolver 2>&1 | tail -2; exit 1; fi
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 > gpurun_out/s38_bench.json 2> gpurun_out/s38_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s38_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['bond_update'])"
timeout -s KILL 100 python tools/sweep_bench.py golden p 24 800 -1 2 > gpurun_out/s38_sweep_golden.json 2> gpurun_out/s38_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s38_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items()})"
timeout -s KILL 200 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 2 > gpurun_out/s38_sweep_hq.json 2> gpurun_out/s38_sweep_hq.err; echo "sweep rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s38_sweep_hq.json'))
for s in d['sweeps']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items()})"
