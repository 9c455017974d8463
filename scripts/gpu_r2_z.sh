#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "bf16 or tf32 or wgrad or pair" 2>&1 | tail -4
for shape in "96 96" "32 32" "128 96" "64 64" "128 128" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $shape --prec bf16 --reps 7 --only wgrad
done
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec tf32 --reps 7 --only wgrad
timeout 120 python scripts/microbench_conv.py 200000 128 128 --prec bf16 --reps 7 --only wgrad
timeout 120 python scripts/microbench_conv.py 30000 256 256 --prec bf16 --reps 7 --only wgrad
} > gpurun_out/r2z6.log 2>&1
cat gpurun_out/r2z6.log
