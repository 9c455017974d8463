#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "conv" 2>&1 | tail -3
echo "=== default (G=1 fwd; wgrad RM auto)"
for c in "96 96" "32 32" "64 64" "128 128" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $c --prec bf16 2>&1 | tail -3
done
echo "=== fwd multi-chunk stages (dbg 4=2), wgrad single visits (5=1)"
for c in "96 96" "64 64" "128 128" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $c --prec bf16 --dbg "4=2,5=1" 2>&1 | tail -3
done
echo "=== wgrad RM=3 (5=2)"
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --only wgrad --dbg "5=2" 2>&1 | tail -1
echo "=== tf32"
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec tf32 2>&1 | tail -3
} > gpurun_out/r2e.log 2>&1
cat gpurun_out/r2e.log
