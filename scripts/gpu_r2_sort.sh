#!/bin/bash
# engine-side row order on the faithful geometry (config 2B): parity test, then the step with / without it
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s --timeout 400 -k "row_order" 2>&1 | tail -8
for scale in 0.34 0.56; do
  for sort in 1 0; do
    SPARSECONV_B200_SORT_ROWS=$sort timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt-precision --geometry faithful --scene-scale $scale > gpurun_out/r2s_2B_${scale}_sort$sort.log 2> gpurun_out/r2s_2B_${scale}_sort$sort.err
    tail -2 gpurun_out/r2s_2B_${scale}_sort$sort.err | cut -c1-300
  done
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-alt-precision > gpurun_out/r2s_dense.log 2> gpurun_out/r2s_dense.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
        print("    ", d["kernel_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
} > gpurun_out/r2s_sort.log 2>&1
cat gpurun_out/r2s_sort.log
