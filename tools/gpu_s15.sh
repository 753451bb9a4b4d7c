#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s15_bench_n4.json 2> gpurun_out/s15_bench_n4.err; echo "n4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/s15_bench_n4.json')); print(d['value'], d['ms_per_step'], d['bond_update'])"
tail -3 gpurun_out/s15_bench_n4.err
