#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s39_bench_n2.json 2> gpurun_out/s39_bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s39_bench_n2.json')); print(d['value'], d['e2e']['value'], d['n_gpus'], d['config']['parallelism'], d['bond_update'])" || tail -20 gpurun_out/s39_bench_n2.err
