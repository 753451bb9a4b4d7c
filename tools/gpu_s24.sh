#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 3 > gpurun_out/s24_sweep_hq_n4.json 2> gpurun_out/s24_sweep_hq_n4.err; echo "rc=$?"
python -c "
import json
txt=[l for l in open('gpurun_out/s24_sweep_hq_n4.json').read().splitlines() if l.startswith('{')][0]
d=json.loads(txt); print(d['n_gpus'])
for s in d['sweeps']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items()})"
tail -3 gpurun_out/s24_sweep_hq_n4.err | cut -c1-300
