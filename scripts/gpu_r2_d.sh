#!/bin/bash
mkdir -p gpurun_out
{
for dbg in "4=1" "4=1,3=1" "4=1,2=1" "3=1" "2=1"; do
  echo "=== fwd dbg=[$dbg]"
  timeout 120 python scripts/microbench_conv.py 1000000 96 96 --prec bf16 --only fwd --dbg "$dbg" 2>&1 | tail -1
done
} > gpurun_out/r2d.log 2>&1
cat gpurun_out/r2d.log
