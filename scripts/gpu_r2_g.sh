#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_unet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r2g_ncu_unet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_resnet.csv python scripts/bench_resnet14.py --batch 16 --steps 3 > gpurun_out/r2g_ncu_resnet.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2g_launches_unet.csv > gpurun_out/r2g_launches_unet.txt 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2g_launches_resnet.csv > gpurun_out/r2g_launches_resnet.txt 2>&1
gzip -f gpurun_out/r2g_launches_unet.csv gpurun_out/r2g_launches_resnet.csv
head -50 gpurun_out/r2g_launches_unet.txt; head -50 gpurun_out/r2g_launches_resnet.txt
