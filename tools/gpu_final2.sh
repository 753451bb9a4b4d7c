#!/bin/bash
# Second (last) short GPU call of round 2: one ncu --set full capture of the two-CTA mix kernel on BASELINE configs[1], then the launch variants.
mkdir -p gpurun_out
O=gpurun_out/final2
timeout ${NCU_TIMEOUT:-55} ncu --set full --clock-control none --import-source on -k regex:mix_smem2 -s 1 -c 1 -o ${O}_ncu_mix2_cfg2 -f python tools/ncu_mix.py cfg2 > ${O}_ncu_mix2_cfg2.out 2>&1
echo "[final2] ncu rc=$?"; tail -2 ${O}_ncu_mix2_cfg2.out
timeout ${AB_TIMEOUT:-70} python tools/ab_mix_variants.py cfg2 cfg4s > ${O}_mix_variants.jsonl 2> ${O}_mix_variants.err
echo "[final2] variants rc=$?"
cat ${O}_mix_variants.jsonl; tail -3 ${O}_mix_variants.err
