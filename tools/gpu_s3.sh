#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q > gpurun_out/s3_pytest_shard.log 2>&1; echo "shard pytest rc=$?"; tail -5 gpurun_out/s3_pytest_shard.log
for ex in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --exchange $ex > gpurun_out/s3_bench_n2_$ex.json 2> gpurun_out/s3_bench_n2_$ex.err
  echo "n2 $ex rc=$?"; cut -c1-330 gpurun_out/s3_bench_n2_$ex.json; tail -3 gpurun_out/s3_bench_n2_$ex.err
done
