"""Where the gradient all-reduce runs relative to the convolution kernels (VERDICT r1 missing-7): a CUPTI kernel
timeline (torch.profiler) of two data-parallel Res16UNet34C steps on rank 0, reduced to a few numbers and a coarse
per-step lane chart.  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/timeline_nccl.py [--voxels 1000000 --scenes 2] > profiles/r2_timeline_8gpu.txt
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth, trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxels", type=int, default=1_000_000)
ap.add_argument("--scenes", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--bucket-mb", type=float, default=25.0)
args = ap.parse_args()
rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    import datetime
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
ops.set_default_precision(args.precision)
torch.manual_seed(0)
model = models.Res16UNet34C(27, 20).to(dev).train()
tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4, bucket_mb=args.bucket_mb)
tr.time_exposed = world > 1
coords, feats, labels = synth.room_batch(777 + rank, args.scenes, args.voxels)
c, f, y = (torch.from_numpy(a).to(dev) for a in (coords, feats, labels))


def step():
    field = ME.TensorField(coordinates=c, features=f)
    loss = ops.cross_entropy(model(field), y, ignore_index=255)
    tr.backward_and_step(loss)


for _ in range(4):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

NSTEPS = 3   # the first profiled step carries the ranks' profiler start-up skew (its first all-reduce waits for the
             # slowest rank to begin): report it, judge by the later ones
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEPS):
        step()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start

    def kind(name):
        n = name.lower()
        if "nccl" in n:
            return "nccl"
        if "conv_umma" in n or "conv_wgrad" in n:
            return "conv"
        if "bn_" in n or "col_reduce" in n:
            return "bn"
        if "sgd" in n:
            return "sgd"
        return "other"

    def union(iv):
        iv = sorted(iv)
        out, cur = 0.0, None
        for a, b in iv:
            if cur is None or a > cur[1]:
                if cur:
                    out += cur[1] - cur[0]
                cur = [a, b]
            else:
                cur[1] = max(cur[1], b)
        if cur:
            out += cur[1] - cur[0]
        return out

    def overlap(a_iv, b_iv):
        return union(a_iv) + union(b_iv) - union(a_iv + b_iv)

    iv = {}
    for e in ev:
        iv.setdefault(kind(e.name), []).append((e.time_range.start - t0, e.time_range.end - t0))
    total = ev[-1].time_range.end - t0
    print(f"# rank 0 of {world}, {NSTEPS} steps (step 0 = profiler start-up skew between ranks), {args.scenes} x {args.voxels} voxels, "
          f"{args.precision}, buckets of {args.bucket_mb} MB")
    print(f"timeline length            {total / 1e3:9.3f} ms")
    for k in ("conv", "bn", "other", "sgd", "nccl"):
        if k in iv:
            print(f"{k:6s} kernels: n = {len(iv[k]):5d}  busy (union) {union(iv[k]) / 1e3:9.3f} ms")
    if "nccl" in iv:
        compute = [x for k in iv if k != "nccl" for x in iv[k]]
        ov = overlap(iv["nccl"], compute)
        nu = union(iv["nccl"])
        print(f"nccl busy {nu / 1e3:.3f} ms, of which {ov / 1e3:.3f} ms ({100 * ov / max(nu, 1e-9):.1f} %) concurrent with a compute kernel")
        # per step: end of the last compute kernel before the SGD kernel vs end of the last nccl kernel
        sg = sorted(iv.get("sgd", []))
        for i, (s0, s1) in enumerate(sg):
            prev_compute_end = max((b for a, b in compute if b <= s0 and (a, b) not in sg), default=0.0)
            last_nccl_end = max((b for a, b in iv["nccl"] if b <= s0 + 1), default=0.0)
            print(f"step {i}: backward's last kernel ends at {prev_compute_end / 1e3:9.3f} ms, last all-reduce kernel at "
                  f"{last_nccl_end / 1e3:9.3f} ms, SGD starts at {s0 / 1e3:9.3f} ms -> tail exposed "
                  f"{max(0.0, s0 - prev_compute_end) / 1e3:.3f} ms")
        prev = 0.0
        for i, (s0, s1) in enumerate(sg):
            st_nccl = [(a, b) for a, b in iv["nccl"] if prev <= a < s1]
            st_comp = [(a, b) for a, b in compute if prev <= a < s1]
            nu_i = union(st_nccl)
            print(f"step {i}: length {(s1 - prev) / 1e3:8.3f} ms, nccl busy {nu_i / 1e3:7.3f} ms, hidden behind compute "
                  f"{overlap(st_nccl, st_comp) / 1e3:7.3f} ms")
            prev = s1
        ex = tr.exposed_allreduce_ms()[-NSTEPS:]
        print("compute-stream wait for the all-reduce handles (CUDA events): " + ", ".join(f"{v:.3f} ms" for v in ex))
        print("nccl kernels (start ms, duration ms):")
        for a, b in sorted(iv["nccl"]):
            print(f"   {a / 1e3:9.3f}  {(b - a) / 1e3:8.3f}")
    # coarse lanes: 100 columns per step pair
    cols = 120
    for k in ("conv", "bn", "other", "nccl"):
        lane = [" "] * cols
        for a, b in iv.get(k, []):
            for x in range(int(a / total * cols), min(cols, int(b / total * cols) + 1)):
                lane[x] = "#"
        print(f"{k:6s}|{''.join(lane)}|")
if world > 1:
    dist.destroy_process_group()
