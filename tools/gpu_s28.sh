#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s28_launches_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s28_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_kernel -c 2 -o gpurun_out/s28_gemm_cfg2_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s28_ncu_full.log 2>&1; echo "ncu full rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/s28_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['bond_update']['total_s'])"
