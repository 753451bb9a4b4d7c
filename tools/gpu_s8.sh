#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/s8_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_kernel -c 2 -o gpurun_out/s8_gemm_cfg2_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s8_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/s8_*
