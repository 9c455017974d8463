#!/bin/bash
# r2t: 2-kernel coordinate insert (128-bit CAS + single-pass row assignment), wgrad with round-robin row blocks
mkdir -p gpurun_out
{
echo "== coordinate / map parity tests"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "coords or insert or stride or kernel_map or unique or quant or pyramid or round_trip or wgrad or conv" 2>&1 | tail -6
echo "== sweep maps"
timeout 300 python scripts/sweep_maps.py 100000 1000000 10000000 2>&1 | grep -v "^$" | head -40
echo "== wgrad"
for shape in "96 96" "32 32" "128 96" "64 64" "128 128" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $shape --prec bf16 --reps 7 --only wgrad
done
timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec tf32 --reps 7 --only wgrad
echo "== wgrad, rows sorted in 64K windows"
for shape in "96 96" "128 96"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $shape --prec bf16 --reps 7 --only wgrad --sort-window 65536
done
} > gpurun_out/r2t.log 2>&1
cat gpurun_out/r2t.log
