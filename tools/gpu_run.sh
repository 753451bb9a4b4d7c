#!/bin/bash
# One parameterised GPU-box script (replaces the 40 one-off tools/gpu_s*.sh of round 1).  Usage, under gpurun:
#   tools/gpu_run.sh TAG STEP [STEP...]      outputs go to gpurun_out/TAG_*
# Steps:
#   tests[:EXPR]        pytest -m gpu (optionally -k EXPR)
#   bench[:WORKLOAD]    python bench.py (N=1) on WORKLOAD (default cfg2)
#   benchq[:WORKLOAD]   the same without cpu baseline / sweep (quick kernel check)
#   mbench:N[:WORKLOAD] torchrun bench.py on N GPUs (no sweep, no cpu baseline)
#   ref[:N]             bench.py --impl reference (under torchrun when N > 1)
#   launches[:WORKLOAD] ncu launch list (gpu__time_duration) of a short bench run
#   ncufull:KERNEL_REGEX[:WORKLOAD]  one ncu --set full capture (first 3 matching launches after 20 skipped), raw + details pages
#   driver:CASE         tools/driver_times.py CASE (cfg1..cfg4 physical runs; the CPU oracle's s/sweep beside the GPU's)
#   sweep:ARGS          tools/sweep_bench.py ARGS (comma-separated -> spaces, ';' -> ',')
#   heig:ARGS           tools/heig_bench.py ARGS
#   py:FILE             python FILE
#   abmix[:OLD.so]      tools/ab_mix.py OLD.so (default build/ab/libchain_b200_head.so) against the product library; then tools/ab_mix_variants.py
#   ncumix[:WORKLOAD]   one ncu --set full capture of the mix kernel (tools/ncu_mix.py, no torch import: starts in seconds)
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
for step in "$@"; do
  IFS=':' read -r kind a b c <<< "$step"
  case $kind in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q ${a:+-k "$a"} > ${O}_pytest.log 2>&1; echo "[$TAG] pytest rc=$?"; tail -5 ${O}_pytest.log ;;
    bench|benchq)
      wl=${a:-cfg2}; extra=""; [ $kind = benchq ] && extra="--no-cpu-baseline --no-sweep"
      timeout 1200 python bench.py --steps 10 --warmup 3 --workload $wl $extra > ${O}_bench_${wl}.json 2> ${O}_bench_${wl}.err; echo "[$TAG] bench $wl rc=$?"
      python tools/show_bench.py ${O}_bench_${wl}.json || tail -20 ${O}_bench_${wl}.err ;;
    mbench)
      n=$a; wl=${b:-cfg2}
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-sweep > ${O}_bench_${wl}_n$n.json 2> ${O}_bench_${wl}_n$n.err; echo "[$TAG] bench $wl N=$n rc=$?"
      python tools/show_bench.py ${O}_bench_${wl}_n$n.json || tail -30 ${O}_bench_${wl}_n$n.err ;;
    ref)
      n=${a:-1}
      if [ "$n" = 1 ]; then timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_ref_n1.json 2> ${O}_ref_n1.err
      else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $n --steps 3 --warmup 1 > ${O}_ref_n$n.json 2> ${O}_ref_n$n.err; fi
      echo "[$TAG] ref N=$n rc=$?"; cat ${O}_ref_n$n.json | cut -c1-600 ;;
    launches)
      wl=${a:-cfg2}
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches_${wl}.csv python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline --no-sweep > ${O}_launches_${wl}.out 2>&1; echo "[$TAG] launches rc=$?"
      python tools/show_launches.py ${O}_launches_${wl}.csv | head -40 ;;
    ncufull)
      # ncufull:KERNEL_REGEX:WORKLOAD:SKIP:COUNT:NAME  (one ncu --set full capture of COUNT launches after SKIP matching ones)
      IFS=':' read -r kind a b c d e <<< "$step"
      wl=${b:-cfg2}; skip=${c:-6}; cnt=${d:-1}; nm=${e:-k}
      timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$a" -s $skip -c $cnt -o ${O}_ncu_${wl}_${nm} -f python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline --no-sweep > ${O}_ncu_${wl}_${nm}.out 2>&1; echo "[$TAG] ncu full rc=$?"
      ncu -i ${O}_ncu_${wl}_${nm}.ncu-rep --page raw --csv > ${O}_ncu_${wl}_${nm}_raw.csv 2>/dev/null
      ncu -i ${O}_ncu_${wl}_${nm}.ncu-rep --page details > ${O}_ncu_${wl}_${nm}_details.txt 2>/dev/null
      ls -la ${O}_ncu_${wl}_${nm}* ;;
    driver)
      timeout 900 python tools/driver_times.py $a > ${O}_driver_${a}.json 2> ${O}_driver_${a}.err; echo "[$TAG] driver $a rc=$?"; cut -c1-900 ${O}_driver_${a}.json ;;
    sweep)
      args=$(echo "$a" | tr ',' ' ' | tr ';' ',')
      timeout 1500 python tools/sweep_bench.py $args > ${O}_sweep.json 2> ${O}_sweep.err; echo "[$TAG] sweep rc=$?"; tail -c 1500 ${O}_sweep.json ;;
    heig)
      args=$(echo "$a" | tr ',' ' ')
      timeout 1200 python tools/heig_bench.py $args > ${O}_heig.log 2>&1; echo "[$TAG] heig rc=$?"; tail -30 ${O}_heig.log ;;
    abmix)
      old=${a:-build/ab/libchain_b200_head.so}
      timeout 300 python tools/ab_mix.py $old chain_b200/libchain_b200.so > ${O}_ab_mix.jsonl 2> ${O}_ab_mix.err; echo "[$TAG] ab_mix rc=$?"; cat ${O}_ab_mix.jsonl
      timeout 300 python tools/ab_mix_variants.py cfg2 cfg4s > ${O}_mix_variants.jsonl 2> ${O}_mix_variants.err; echo "[$TAG] variants rc=$?"; cat ${O}_mix_variants.jsonl ;;
    ncumix)
      wl=${a:-cfg2}
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:mix_smem2 -s 1 -c 1 -o ${O}_ncu_mix2_${wl} -f python tools/ncu_mix.py $wl > ${O}_ncu_mix2_${wl}.out 2>&1; echo "[$TAG] ncu rc=$?"
      ncu -i ${O}_ncu_mix2_${wl}.ncu-rep --page raw --csv > ${O}_ncu_mix2_${wl}_raw.csv 2>/dev/null
      ncu -i ${O}_ncu_mix2_${wl}.ncu-rep --page details > ${O}_ncu_mix2_${wl}_details.txt 2>/dev/null
      ncu -i ${O}_ncu_mix2_${wl}.ncu-rep --page source --csv > ${O}_ncu_mix2_${wl}_source.csv 2>/dev/null ;;
    py)
      timeout 1500 python $a > ${O}_py.log 2>&1; echo "[$TAG] py $a rc=$?"; tail -40 ${O}_py.log ;;
    *) echo "unknown step $step" ;;
  esac
done
