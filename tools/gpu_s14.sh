#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/s14_pytest_k.log 2>&1; rc=$?; echo "gemm pytest rc=$rc"; tail -4 gpurun_out/s14_pytest_k.log
if [ $rc -ne 0 ]; then exit 1; fi
for w in cfg2 cfg5 cfg3; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/s14_bench_$w.json 2> gpurun_out/s14_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s14_bench_$w.json")); r=d["roofline"]
    print("$w value=%.2f ms=%.2f frac=%.3f stage_ms=%s trunc=%.3f"%(d["value"],d["ms_per_step"],r["frac"],{k:round(x,2) for k,x in r["stage_ms"].items()}, d["bond_update"]["truncation_s"]))
except Exception as e: print("$w ERR", e, open("gpurun_out/s14_bench_$w.err").read()[-400:])
PY
done
timeout 200 python tools/calibrate_fp64.py 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s14_pytest.log
