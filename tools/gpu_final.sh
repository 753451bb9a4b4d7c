#!/bin/bash
# Last GPU call of round 2 (a few minutes of box time left): A/B of the two-CTA mix kernel against the previous build, then the GPU test
# suite on the new library, then a short bench line if time remains.  Everything is written to gpurun_out/ as it is produced.
mkdir -p gpurun_out
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > ${O}_gpu.txt 2>&1
timeout ${AB_TIMEOUT:-120} python tools/ab_mix.py build/ab/libchain_b200_head.so chain_b200/libchain_b200.so > ${O}_ab_mix.jsonl 2> ${O}_ab_mix.err
echo "[final] ab_mix rc=$?"
cat ${O}_ab_mix.jsonl
tail -3 ${O}_ab_mix.err
PYTHONUNBUFFERED=1 timeout ${TEST_TIMEOUT:-260} python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_adapter.py tests/test_gpu_kernels.py \
  tests/test_gpu_bench_parity.py > ${O}_pytest.log 2>&1
echo "[final] pytest rc=$?"
tail -15 ${O}_pytest.log
timeout ${BENCH_TIMEOUT:-150} python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err
echo "[final] bench rc=$?"
python tools/show_bench.py ${O}_bench_cfg2.json 2>/dev/null || tail -5 ${O}_bench_cfg2.err
