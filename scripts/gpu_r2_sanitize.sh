#!/bin/bash
# compute-sanitizer memcheck over the kernels written in the last session of round 2 (small inputs)
mkdir -p gpurun_out
{
echo "== memcheck: coordinate insert / maps (pytest subset)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 500 -k "coords_insert or unique or pyramid or stride" 2>&1 | tail -8
echo "== memcheck: CTA-pair kernels + wgrad (pair_check, 60 K voxels)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/exp/pair_check.py 60000 64 96 2>&1 | tail -10
} > gpurun_out/r2_sanitize.log 2>&1
cat gpurun_out/r2_sanitize.log
