#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "conv" 2>&1 | tail -15
for c in "96 96" "32 32" "64 64" "128 128" "128 96" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $c --prec bf16 2>&1 | tail -3
done
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec tf32 2>&1 | tail -3
echo "--- old-style single chunk stages / single visits"
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --dbg "4=1,5=1" 2>&1 | tail -3
timeout 120 python scripts/microbench_conv.py 1000000 128 128 --prec bf16 --dbg "4=1,5=1" 2>&1 | tail -3
echo "--- producer groups"
for g in 1 2; do timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --only fwd --dbg "0=$g" 2>&1 | tail -1; done
for g in 1 2 4; do timeout 120 python scripts/microbench_conv.py 1000000 128 128 --prec bf16 --only fwd --dbg "0=$g" 2>&1 | tail -1; done
echo "--- small maps"
timeout 120 python scripts/microbench_conv.py 200000 96 96 --prec bf16 2>&1 | tail -3
timeout 120 python scripts/microbench_conv.py 8000 256 256 --prec bf16 2>&1 | tail -3
} > gpurun_out/r2c_conv_v2.log 2>&1
cat gpurun_out/r2c_conv_v2.log
