#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "eigensolver" > gpurun_out/s36_pytest_heig.log 2>&1; rc=$?; echo "pytest heig rc=$rc"; tail -3 gpurun_out/s36_pytest_heig.log
if [ $rc -ne 0 ]; then CB200_INVIT=ref timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "eigensolver" > gpurun_out/s36_pytest_heig_ref.log 2>&1; echo "with CB200_INVIT=ref rc=$?"; tail -3 gpurun_out/s36_pytest_heig_ref.log; CB200_TRI_BATCH=0 timeout -s KILL 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "eigensolver" > gpurun_out/s36_pytest_heig_nobatch.log 2>&1; echo "with CB200_TRI_BATCH=0 rc=$?"; tail -3 gpurun_out/s36_pytest_heig_nobatch.log; fi
if [ $rc -ne 0 ]; then tail -40 gpurun_out/s36_pytest_heig.log; exit 1; fi
timeout -s KILL 400 python -m pytest tests -m gpu -x -q > gpurun_out/s36_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s36_pytest.log
timeout -s KILL 300 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 3 > gpurun_out/s36_sweep_hq.json 2> gpurun_out/s36_sweep_hq.err; echo "sweep rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s36_sweep_hq.json'))
for s in d['sweeps']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items()})"
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s36_bench.json 2> gpurun_out/s36_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s36_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['bond_update'])"
for c in cfg2; do timeout -s KILL 200 python tools/driver_times.py $c 2>&1 | tail -1 | cut -c1-600; done | tee gpurun_out/s36_driver_times.log
