#!/bin/bash
# 2-GPU check of the bench line with the CTA-pair forward kernel under NCCL overlap
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-alt-precision > gpurun_out/r2g_bench_n2.log 2> gpurun_out/r2g_bench_n2.err
tail -c 1500 gpurun_out/r2g_bench_n2.log; tail -3 gpurun_out/r2g_bench_n2.err
SPARSECONV_B200_DEBUG_SET="8=1" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-alt-precision > gpurun_out/r2g_bench_n2_nopair.log 2> gpurun_out/r2g_bench_n2_nopair.err
python - <<'PY'
import json
for tag in ("r2g_bench_n2", "r2g_bench_n2_nopair"):
    try:
        d = json.loads(open(f"gpurun_out/{tag}.log").read().strip().splitlines()[-1])
        print(tag, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "n_gpus", d["n_gpus"])
    except Exception as e:
        print(tag, "failed", e)
PY
