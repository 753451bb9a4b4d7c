#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # N exchange
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2961$1 bench.py --gpus $1 --steps 5 --warmup 3 --no-cpu-baseline --exchange $2 > gpurun_out/s9_bench_n$1_$2.json 2> gpurun_out/s9_bench_n$1_$2.err
  echo "n$1 $2 rc=$? $(cut -c1-160 gpurun_out/s9_bench_n$1_$2.json)"
}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_n1.json 2> gpurun_out/s9_bench_n1.err; echo "n1 $(cut -c1-120 gpurun_out/s9_bench_n1.json)"
run 2 p2p; run 4 p2p; run 8 p2p; run 8 nccl; run 4 nccl
