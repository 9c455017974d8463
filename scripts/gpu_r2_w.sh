#!/bin/bash
# r2w: CTA-pair forward kernel: parity against the single-CTA kernel, then timings
mkdir -p gpurun_out
{
timeout 120 python scripts/exp/pair_check.py 200000 96 96 2>&1 | tail -12
echo "rc=$?"
timeout 120 python scripts/exp/pair_check.py 1000000 96 96 --time 2>&1 | tail -14
timeout 120 python scripts/exp/pair_check.py 1000000 32 32 --time 2>&1 | tail -10
timeout 120 python scripts/exp/pair_check.py 1000000 128 128 --time 2>&1 | tail -10
timeout 120 python scripts/exp/pair_check.py 1000000 64 64 --time 2>&1 | tail -10
timeout 120 python scripts/exp/pair_check.py 1000000 128 96 --time 2>&1 | tail -10
timeout 120 python scripts/exp/pair_check.py 1000000 256 256 --time 2>&1 | tail -10
} > gpurun_out/r2w.log 2>&1
cat gpurun_out/r2w.log
