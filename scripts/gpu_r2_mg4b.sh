#!/bin/bash
# 4-GPU: does limiting NCCL's channel count (SMs taken from the persistent convolution kernels) help the step?
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 100)) bench.py --gpus 4 --steps 12 --warmup 3 --no-alt-precision > gpurun_out/r2h_n4_$tag.log 2> gpurun_out/r2h_n4_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2h_n4_{tag}.log").read().strip().splitlines()[-1])
    print(tag, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print(tag, "failed", e)
PY
}
run default A=1
run ch8 NCCL_MAX_NCHANNELS=8
run ch4 NCCL_MAX_NCHANNELS=4
run ch2 NCCL_MAX_NCHANNELS=2
run default2 A=1
