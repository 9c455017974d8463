#!/bin/bash
mkdir -p gpurun_out
{
for dbg in "" "4=1" "5=1" "6=2" "6=4" "4=1,5=1" "5=1,6=2"; do
  echo "=== wgrad dbg=[$dbg]"
  timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --only wgrad --dbg "$dbg" 2>&1 | tail -2
done
for dbg in "" "3=1"; do
  echo "=== fwd dbg=[$dbg]"
  timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --only fwd --dbg "$dbg" 2>&1 | tail -1
done
for c in "64 64" "128 128" "32 32" "256 256"; do
  timeout 120 python scripts/microbench_conv.py 1000000 $c --prec bf16 2>&1 | tail -4
done
} > gpurun_out/r2b_wgrad_dbg.log 2>&1
cat gpurun_out/r2b_wgrad_dbg.log
