#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/s26_pytest_k.log 2>&1; rc=$?; echo "gemm pytest rc=$rc"; tail -4 gpurun_out/s26_pytest_k.log
if [ $rc -ne 0 ]; then exit 1; fi
for m3 in 96 64; do
  CB200_CPLX_3M=$m3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg2 > gpurun_out/s26_bench_cfg2_m3_$m3.json 2> gpurun_out/s26_bench_cfg2_m3_$m3.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s26_bench_cfg2_m3_$m3.json")); r=d["roofline"]
    print("3M=$m3 cfg2 value=%.2f ms=%.2f e2e=%.2f frac=%.3f stage_ms=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["frac"],{k:round(x,2) for k,x in r["stage_ms"].items()}))
except Exception as e: print("3M=$m3 ERR", e, open("gpurun_out/s26_bench_cfg2_m3_$m3.err").read()[-400:])
PY
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s26_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s26_pytest.log
