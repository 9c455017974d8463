#!/bin/bash
# tests + bf16 bench with per-layer detail
mkdir -p gpurun_out
TAG=${TAG:-r1d}
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --detail ${BENCH_ARGS} > gpurun_out/${TAG}_bench_bf16.log 2>&1
grep '^{' gpurun_out/${TAG}_bench_bf16.log | cut -c1-3000
