#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shard.py -x -q > gpurun_out/s27_pytest_shard.log 2>&1; echo "shard pytest rc=$?"; tail -4 gpurun_out/s27_pytest_shard.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s27_bench_n2.json 2> gpurun_out/s27_bench_n2.err; echo "n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/s27_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['bond_update'].items() if k!='note'})"
