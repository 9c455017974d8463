#!/bin/bash
# r2x: CTA-pair wgrad kernel: parity + timings; pair parity tests
mkdir -p gpurun_out
{
timeout 120 python scripts/exp/pair_check.py 200000 96 96 2>&1 | grep -v "bit-identical True\|kernel done\|stats  " | tail -6
for shape in "96 96" "32 32" "128 128" "64 64" "128 96" "256 256"; do
  timeout 120 python scripts/exp/pair_check.py 1000000 $shape --time 2>&1 | grep "wgrad\|WGRAD\|MISMATCH\|rror" | tail -6
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "pair or bf16" 2>&1 | tail -5
} > gpurun_out/r2x.log 2>&1
cat gpurun_out/r2x.log
