#!/bin/bash
# round 2, call A: measure what round 1 built blind + the precision study + LDGSTS microbench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
( cd scripts/exp && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_bench ldgsts_bench.cu 2>&1 | tail -2; timeout 120 ./ldgsts_bench ) > gpurun_out/r2a_ldgsts.log 2>&1
timeout 200 python scripts/measure_tf32_peak.py gpurun_out/r2a_tf32_peak.json > gpurun_out/r2a_tf32_peak.log 2>&1
timeout 600 python scripts/diag_precision_scale.py 200000 2 > gpurun_out/r2a_precision_200k.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --geometry faithful --scene-scale 0.34 --detail > gpurun_out/r2a_bench_2B_034.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --geometry faithful --scene-scale 0.56 --detail > gpurun_out/r2a_bench_2B_056.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --fused-head > gpurun_out/r2a_bench_fused_head.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 16 --steps 20 > gpurun_out/r2a_resnet14_b16.log 2>&1
tail -3 gpurun_out/r2a_*.log
