#!/bin/bash
# Where the forward kernel's time goes (experiment library, -DSPC_EXPERIMENTS): knob 5 = 1 no gather copies,
# 2 no MMAs, 4 no weight slabs (combinable).  usage: conv_breakdown.sh [VOXELS]
V=${1:-1000000}
export SPARSECONV_B200_LIB=$(dirname $0)/../../nerf_downstream_b200/libsparseconv_b200_exp.so
for shape in "96 96" "32 32" "128 128" "256 256"; do
  for dbg in 0 1 2 3 4 5 6 7; do
    echo -n "dbg5=$dbg  "
    python $(dirname $0)/../microbench_conv.py $V $shape --prec bf16 --only fwd --reps 5 --dbg 5=$dbg | tail -1
  done
done
