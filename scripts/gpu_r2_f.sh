#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -25 > gpurun_out/r2f_tests.log
{
for c in "96 96" "128 128" "256 256" "32 32"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $c --prec bf16 --only wgrad 2>&1 | tail -1
done
timeout 300 python scripts/sweep_maps.py 1000000 2>&1 | tail -30
} > gpurun_out/r2f_micro.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --detail > gpurun_out/r2f_bench.log 2>&1
tail -5 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_micro.log; tail -c 6000 gpurun_out/r2f_bench.log
