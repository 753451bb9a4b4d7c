#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q > gpurun_out/s11_pytest_shard.log 2>&1; echo "shard pytest rc=$?"; tail -12 gpurun_out/s11_pytest_shard.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench_n2.json 2> gpurun_out/s11_bench_n2.err; echo "n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s11_bench_n2.json')); print(d['value'], d['bond_update'])"
for c in cfg1 cfg3 cfg2; do timeout 900 python tools/driver_times.py $c 2>&1 | tail -1 | cut -c1-600; done | tee gpurun_out/s11_driver_times.log
timeout 900 python tools/sweep_bench.py golden p 24 800 -1 2 > gpurun_out/s11_sweep_golden.json 2> gpurun_out/s11_sweep_golden.err; cut -c1-1500 gpurun_out/s11_sweep_golden.json
timeout 1500 python tools/sweep_bench.py haagerupq p 12 1600 0,-1,0 2 > gpurun_out/s11_sweep_hq.json 2> gpurun_out/s11_sweep_hq.err; cut -c1-1500 gpurun_out/s11_sweep_hq.json; tail -3 gpurun_out/s11_sweep_hq.err
