#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/s32_pytest_k.log 2>&1; rc=$?; echo "gemm pytest rc=$rc"; tail -2 gpurun_out/s32_pytest_k.log
if [ $rc -ne 0 ]; then exit 1; fi
for v in "" _s3; do
  CHAIN_B200_LIB=chain_b200/libchain_b200$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg2 > gpurun_out/s32_bench$v.json 2> gpurun_out/s32_bench$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s32_bench$v.json")); r=d["roofline"]
    print("lib$v value=%.2f ms=%.2f frac=%.3f stage_ms=%s"%(d["value"],d["ms_per_step"],r["frac"],{k:round(x,2) for k,x in r["stage_ms"].items()}))
except Exception as e: print("lib$v ERR", e)
PY
done
