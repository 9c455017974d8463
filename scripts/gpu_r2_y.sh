#!/bin/bash
# r2y: whole GPU suite + the step with the CTA-pair forward kernel
mkdir -p gpurun_out
{
echo "== GPU test suite"
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6
echo "== bench (default)"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2y_bench.log 2> gpurun_out/r2y_bench.err; tail -3 gpurun_out/r2y_bench.err
echo "== bench, pair kernel off (knob 8 = 1)"
SPARSECONV_B200_DEBUG_SET="8=1" timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-precision --no-cpu-baseline > gpurun_out/r2y_bench_nopair.log 2> gpurun_out/r2y_bench_nopair.err
echo "== bench, pair kernel on, second run"
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-precision --no-cpu-baseline > gpurun_out/r2y_bench2.log 2> gpurun_out/r2y_bench2.err
python - <<'PY'
import json
for tag in ("bench", "bench_nopair", "bench2"):
    try:
        d = json.loads(open(f"gpurun_out/r2y_{tag}.log").read().strip().splitlines()[-1])
        print(tag, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "launches", d.get("gpu_launches"),
              "roofline", d["roofline"]["kernel"], "%.3f" % d["roofline"]["frac"])
        print("   classes", d.get("kernel_classes_ms_per_step"))
        if "alt_precision" in d: print("   tf32 %.4g" % d["alt_precision"]["value"], "ms %.2f" % d["alt_precision"]["ms_per_step"])
    except Exception as e:
        print(tag, "failed", e)
PY
} > gpurun_out/r2y.log 2>&1
cat gpurun_out/r2y.log
