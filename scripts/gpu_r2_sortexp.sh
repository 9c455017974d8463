#!/bin/bash
# effect of mask-sorted row order on the convolution kernels (same voxels, re-inserted in sorted order)
mkdir -p gpurun_out
{
for W in 0 16384 65536 1000000; do
  for shape in "96 96" "32 32" "128 96" "64 64"; do
    echo "== sort-window $W  $shape"
    python scripts/microbench_conv.py 1000000 $shape --prec bf16 --reps 7 --sort-window $W
  done
done
} > gpurun_out/r2s_sortexp.log 2>&1
tail -60 gpurun_out/r2s_sortexp.log
