#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 scripts/bench_resnet14.py --batch 16 --steps 30 --json gpurun_out/r2_config3_n24.jsonl > gpurun_out/r2mg4_resnet14_n2.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29542 scripts/bench_resnet14.py --batch 16 --steps 30 --json gpurun_out/r2_config3_n24.jsonl > gpurun_out/r2mg4_resnet14_n4.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 16 --steps 30 --json gpurun_out/r2_config3_n24.jsonl > gpurun_out/r2mg4_resnet14_n1.log 2>&1
timeout 400 $TR --nproc-per-node 4 --master-port 29543 scripts/timeline_nccl.py > gpurun_out/r2mg4_timeline_unet_n4.txt 2> gpurun_out/r2mg4_timeline_unet_n4.err
cat gpurun_out/r2_config3_n24.jsonl; head -24 gpurun_out/r2mg4_timeline_unet_n4.txt
