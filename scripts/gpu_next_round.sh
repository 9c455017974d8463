#!/bin/bash
# First GPU call of the next round: everything that was built without GPU time left.
#   - full GPU suite + smoke (incl. tests/test_gpu_widen.py, tests/test_gpu_zbackbones.py)
#   - default bench, bench --fused-head, bench --geometry faithful (config 2B, scene_scale 0.34 and 0.56)
#   - ncu --set full of the widened kernels (seg head, instance norm, interpolation): dram bytes for their roofline lines
# Logs land in gpurun_out/; budget ~6 GPU-minutes.   usage: gpurun --timeout 900 -- 'bash scripts/gpu_next_round.sh'
mkdir -p gpurun_out
TAG=${TAG:-r2a}
timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for args in "" "--fused-head" "--geometry faithful --scene-scale 0.34" "--geometry faithful --scene-scale 0.56"; do
  name=$(echo "bench${args}" | tr ' ' '_' | tr -d '-')
  timeout 600 python bench.py --no-cpu-baseline --detail $args > gpurun_out/${TAG}_${name}.log 2>&1
  grep '^{' gpurun_out/${TAG}_${name}.log | cut -c1-600
done
timeout 120 python scripts/bench_widen.py > gpurun_out/${TAG}_bench_widen.log 2>&1; cat gpurun_out/${TAG}_bench_widen.log
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'seg_head_kernel|inst_sums4_kernel|inst_apply4_kernel|inst_bwd_apply4_kernel|interp_fwd_kernel|interp_bwd_kernel' -c 12 \
   -o gpurun_out/prof_${TAG}_widen -f python scripts/bench_widen.py > gpurun_out/${TAG}_ncu_widen.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_widen.log
