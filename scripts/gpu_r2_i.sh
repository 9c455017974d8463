#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -12 > gpurun_out/r2i_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --geometry faithful --scene-scale 0.34 > gpurun_out/r2i_bench_2B_034.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --geometry faithful --scene-scale 0.56 > gpurun_out/r2i_bench_2B_056.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --fused-head > gpurun_out/r2i_bench_fused_head.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --scenes 1 > gpurun_out/r2i_bench_1scene.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-voxels 1000000 --cpu-forward-only > gpurun_out/r2i_cpu_same_scene_fwd.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_cpu_default.log 2>&1
timeout 400 python scripts/bench_resnet14.py --batch 4 --steps 20 --cpu > gpurun_out/r2i_resnet14_b4_cpu.log 2>&1
tail -4 gpurun_out/r2i_tests.log
for f in gpurun_out/r2i_bench*.log gpurun_out/r2i_cpu*.log; do echo "== $f"; grep -E '^\{' $f | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('impl','ours'), d['dtype'], round(d['value']/1e6,3),'Mvox/s', round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['value']/1e6,3), d.get('gpu_launches'), d['config'].get('sample_voxels'), (d.get('alt_precision') or {}).get('value'))
"; done
tail -3 gpurun_out/r2i_resnet14_b4_cpu.log
