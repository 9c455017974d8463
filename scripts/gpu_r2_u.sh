#!/bin/bash
# r2u: z-pair home buckets; per-kernel durations of the coordinate kernels
mkdir -p gpurun_out
{
echo "== coordinate / map parity tests"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "coords or insert or stride or kernel_map or unique or quant or pyramid or round_trip or symmetric or interp" 2>&1 | tail -4
echo "== sweep maps"
timeout 300 python scripts/sweep_maps.py 1000000 10000000 2>&1 | grep -v "^$" | grep -v "BN\|ReLU"
echo "== sweep maps, shuffled voxel order (1 M)"
SWEEP_SHUFFLE=1 timeout 300 python scripts/sweep_maps.py 1000000 2>&1 | grep -v "^$" | grep -v "BN\|ReLU" | head -8
echo "== per-kernel durations (ncu, serialised)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches.csv python scripts/sweep_maps.py 1000000 > /dev/null 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2u_launches.csv
} > gpurun_out/r2u.log 2>&1
cat gpurun_out/r2u.log
