#!/bin/bash
# evidence for profiles/: launch list of the default bench, full ncu capture of the dominant conv kernels,
# config-4 (hash / kernel-map sweep) and config-5 (layer microbench sweep) logs
mkdir -p gpurun_out
TAG=${TAG:-r1final}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2600 --csv \
   --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -c 300 gpurun_out/${TAG}_ncu_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_umma_kernel|conv_wgrad_umma_kernel' -s 1 -c 5 \
   -o gpurun_out/prof_${TAG}_conv96 -f python scripts/microbench_conv.py 1000000 96 96 --reps 1 --prec bf16 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
timeout 600 python scripts/sweep_maps.py 100000 300000 1000000 3000000 10000000 > gpurun_out/${TAG}_sweep_maps.log 2>&1
for prec in bf16 tf32; do
  for cfg in "1000000 32 32" "1000000 64 64" "1000000 96 96" "1000000 128 128" "1000000 256 256" "1000000 128 96" "1000000 96 96 2 2"; do
    timeout 200 python scripts/microbench_conv.py $cfg --prec $prec 2>&1 | grep -E "^(fwd|dgrad|wgrad) "
  done
done > gpurun_out/${TAG}_sweep_layers.log 2>&1
tail -4 gpurun_out/${TAG}_sweep_layers.log
