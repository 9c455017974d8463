#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -25 > gpurun_out/r2h_tests.log
timeout 300 python scripts/sweep_maps.py 1000000 2>&1 | grep -E "kernel map 3\^3 s1|quantize|tile mask|transpose" > gpurun_out/r2h_maps.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 16 --steps 20 > gpurun_out/r2h_resnet14.log 2>&1
timeout 300 python scripts/bench_resnet14.py --batch 4 --steps 20 >> gpurun_out/r2h_resnet14.log 2>&1
tail -8 gpurun_out/r2h_tests.log; cat gpurun_out/r2h_maps.log gpurun_out/r2h_resnet14.log; tail -c 5000 gpurun_out/r2h_bench.log
