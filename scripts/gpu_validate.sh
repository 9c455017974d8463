#!/bin/bash
# full validation: all GPU tests, smoke, default bench (bf16, 2 scenes), tf32 bench, reference arm
mkdir -p gpurun_out
TAG=${TAG:-r1i}
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --detail > gpurun_out/${TAG}_bench_default.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_default.log | cut -c1-1200
timeout 600 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/${TAG}_bench_tf32.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_tf32.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1; grep '^{' gpurun_out/${TAG}_bench_reference.log | cut -c1-400
