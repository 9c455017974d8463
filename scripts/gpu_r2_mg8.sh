#!/bin/bash
# 8-GPU measurements: config 3 (ResNet14 DP, B=16/GPU) and the NCCL/compute timeline of the UNet step
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2mg8_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 scripts/bench_resnet14.py --batch 16 --steps 30 --json gpurun_out/r2_config3.jsonl > gpurun_out/r2mg8_resnet14_n8.log 2>&1
timeout 400 $TR --nproc-per-node 8 --master-port 29532 scripts/timeline_nccl.py > gpurun_out/r2mg8_timeline_unet_n8.txt 2> gpurun_out/r2mg8_timeline_unet_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29533 scripts/bench_resnet14.py --batch 16 --steps 30 --precision tf32 --json gpurun_out/r2_config3.jsonl > gpurun_out/r2mg8_resnet14_n8_tf32.log 2>&1
tail -2 gpurun_out/r2mg8_resnet14_n8.log gpurun_out/r2mg8_resnet14_n8_tf32.log; head -30 gpurun_out/r2mg8_timeline_unet_n8.txt; tail -3 gpurun_out/r2mg8_timeline_unet_n8.err
