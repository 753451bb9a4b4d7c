#!/bin/bash
# session-1 GPU pass: parity tests, Jacobi inner-sweep experiment, bench (both arms), launch list of one bond update
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" 
for inner in 0 1 2 3; do
  for w in cfg3 tiny; do
    echo "inner=$inner $w: $(CB200_JACOBI_INNER=$inner python tools/bond_update_times.py $w 2>&1 | tail -1)" >> gpurun_out/s1_jacobi.log
  done
done
CB200_JACOBI_INNER=2 python tools/bond_update_times.py cfg2 >> gpurun_out/s1_jacobi.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s1_bench_ref.json 2> gpurun_out/s1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s1_launches_update_cfg3.csv python tools/bond_update_times.py cfg3 > gpurun_out/s1_ncu_update.log 2>&1
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_jacobi.log; cat gpurun_out/s1_bench.json | cut -c1-600
