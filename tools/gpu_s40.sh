#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 110 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s40_bench.json 2> gpurun_out/s40_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s40_bench.json')); print(d['value'], d['e2e']['value'], d['bond_update']['total_s'], d['sweep'])" || tail -20 gpurun_out/s40_bench.err
