#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s20_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/s20_bench_cfg4.json 2> gpurun_out/s20_bench_cfg4.err; echo "cfg4 rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s20_bench_cfg4.json")); r=d["roofline"]
    print("cfg4 value=%.2f ms=%.2f e2e=%.2f frac=%.3f bond=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["frac"], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["bond_update"].items() if k!="note"}))
except Exception as e: print("cfg4 ERR", e, open("gpurun_out/s20_bench_cfg4.err").read()[-600:])
PY
timeout 900 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 3 > gpurun_out/s20_sweep_hq.json 2> gpurun_out/s20_sweep_hq.err; python -c "
import json; d=json.load(open('gpurun_out/s20_sweep_hq.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
timeout 600 python tools/sweep_bench.py golden p 24 800 -1 3 > gpurun_out/s20_sweep_golden.json 2> gpurun_out/s20_sweep_golden.err; python -c "
import json; d=json.load(open('gpurun_out/s20_sweep_golden.json'))
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
