"""BASELINE.json configs[0] / configs[2] (parity-test cases, not the headline): co3d_3d ResNet14(27 -> 51) sparse
classifier fwd + bwd + SGD on a synthetic CO3D-shaped plenoxel batch, B objects per GPU (reference: 16,
co3d_cls.gin:28).  Prints voxels/s and objects/s; `--cpu` also times the CPU restatement on a B=4 batch.
Under `torchrun --nproc-per-node N` it is SURVEY.md §8d config 3: B objects per GPU (ranks seeded 777 + rank), weak
scaling, gradient all-reduce over NCCL, time = max over ranks.
usage: bench_resnet14.py [--batch 16] [--steps 20] [--precision bf16|tf32|fp32] [--cpu]"""
import argparse
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth, trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--json", default="", help="append a JSON line with the result to this file (rank 0)")
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
model = models.ResNet14(27, 51).to(dev).train()
tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4)
tr.time_exposed = world > 1   # device time the compute stream waits for gradient all-reduces after backward
coords, feats, labels = synth.co3d_batch(777 + rank, args.batch)
c, f, y = (torch.from_numpy(a).to(dev) for a in (coords, feats, labels))


def step():
    field = ME.TensorField(coordinates=c, features=f)
    logits = model(field)
    loss = ops.cross_entropy(logits, y)
    tr.backward_and_step(loss)
    return field.coordinate_manager.size(field.coordinate_manager.get_unique_coordinate_map_key(1))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(5):
    vox = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
barrier()
exposed = tr.exposed_allreduce_ms()[-args.steps:] if world > 1 else []
exposed_ms = sum(exposed) / max(len(exposed), 1)
stats = torch.tensor([e0.elapsed_time(e1) / args.steps, float(vox), exposed_ms, float(vox), float(vox)], device=dev,
                     dtype=torch.float64)
if world > 1:
    t = stats[:1].clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # time = the slowest rank
    v = stats[1:2].clone()
    dist.all_reduce(v, op=dist.ReduceOp.SUM)            # voxels of the whole job
    ex = stats[2:3].clone()
    dist.all_reduce(ex, op=dist.ReduceOp.MAX)           # exposed all-reduce: the worst rank
    vmax, vmin = stats[3:4].clone(), stats[4:5].clone()
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)         # per-rank voxel imbalance (ranks draw different objects)
    dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
    stats = torch.cat([t, v, ex, vmax, vmin])
ms, vox, exposed_ms, vmax, vmin = float(stats[0]), int(stats[1]), float(stats[2]), int(stats[3]), int(stats[4])
if rank == 0:
    imb = vmax / (vox / world)
    print(f"ResNet14 fwd+bwd+SGD  {world} GPU(s)  B={args.batch}/GPU  {vox} voxels/step  {args.precision}: {ms:.3f} ms/step  "
          f"{vox / ms / 1e3:.2f} M voxels/s  {world * args.batch / ms * 1e3:.0f} objects/s  "
          f"exposed all-reduce {exposed_ms:.3f} ms/step (max over ranks)  voxels per rank min {vmin} max {vmax} "
          f"(max / mean = {imb:.3f})", flush=True)
    if args.json:
        import json
        with open(args.json, "a") as fh:
            fh.write(json.dumps({"config": "BASELINE configs[2]: ResNet14(27->51) data parallel, weak scaling",
                                 "n_gpus": world, "batch_per_gpu": args.batch, "precision": args.precision,
                                 "voxels_per_step": vox, "ms_per_step": ms, "voxels_per_s": vox / ms * 1e3,
                                 "objects_per_s": world * args.batch / ms * 1e3, "exposed_allreduce_ms_per_step": exposed_ms,
                                 "voxels_per_rank_min": vmin, "voxels_per_rank_max": vmax,
                                 "imbalance_max_over_mean": imb, "gradient_bytes": 4 * tr.arena.numel}) + "\n")
if world > 1:
    dist.destroy_process_group()
if args.cpu and rank == 0:
    from oracle import nets
    cc, ff, yy = synth.co3d_batch(777, 4)
    params = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in model.state_dict().items()}
    ft, yt = torch.from_numpy(ff), torch.from_numpy(yy)

    def cpu_step():
        for p in params.values():
            p.grad = None
        loss = torch.nn.functional.cross_entropy(nets.resnet_forward(params, cc, ft, use_c=True), yt)
        loss.backward()
    cpu_step()
    t0 = time.perf_counter()
    for _ in range(3):
        cpu_step()
    dt = (time.perf_counter() - t0) / 3
    print(f"CPU restatement of ME's algorithm (ME not installable), B=4, {torch.get_num_threads()} threads: "
          f"{dt * 1e3:.1f} ms/step  {4 / dt:.1f} objects/s", flush=True)
