#!/bin/bash
mkdir -p gpurun_out
for v in "" _v1; do
  for w in cfg2 cfg3 cfg5; do
    CHAIN_B200_LIB=chain_b200/libchain_b200$v.so timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/s13_bench_${w}$v.json 2> gpurun_out/s13_bench_${w}$v.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s13_bench_${w}$v.json")); r=d["roofline"]
    print("lib$v $w value=%.2f ms=%.2f frac=%.3f stage_ms=%s trunc=%.3f"%(d["value"],d["ms_per_step"],r["frac"],{k:round(x,2) for k,x in r["stage_ms"].items()}, d["bond_update"]["truncation_s"]))
except Exception as e: print("lib$v $w ERR", e)
PY
  done
done
CHAIN_B200_LIB=chain_b200/libchain_b200_v1.so python tools/calibrate_fp64.py 2>&1 | tail -12
