#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/s18_bench_cfg4.json 2> gpurun_out/s18_bench_cfg4.err; echo "cfg4 rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s18_bench_cfg4.json")); r=d["roofline"]
    print("cfg4 value=%.2f ms=%.2f e2e=%.2f frac=%.3f stage_ms=%s bond=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["frac"],{k:round(x,2) for k,x in r["stage_ms"].items()}, {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["bond_update"].items() if k!="note"}))
except Exception as e: print("cfg4 ERR", e, open("gpurun_out/s18_bench_cfg4.err").read()[-600:])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
