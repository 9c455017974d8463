"""BASELINE.json configs[0] / configs[2] (parity-test cases, not the headline): co3d_3d ResNet14(27 -> 51) sparse
classifier fwd + bwd + SGD on a synthetic CO3D-shaped plenoxel batch, B objects per GPU (reference: 16,
co3d_cls.gin:28).  Prints voxels/s and objects/s; `--cpu` also times the CPU restatement on a B=4 batch.
usage: bench_resnet14.py [--batch 16] [--steps 20] [--precision bf16|tf32|fp32] [--cpu]"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth, trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
ops.set_default_precision(args.precision)
torch.manual_seed(0)
model = models.ResNet14(27, 51).to(dev).train()
tr = trainer.DataParallelTrainer(model, lr=0.1, momentum=0.9, weight_decay=1e-4)
coords, feats, labels = synth.co3d_batch(777, args.batch)
c, f, y = (torch.from_numpy(a).to(dev) for a in (coords, feats, labels))


def step():
    field = ME.TensorField(coordinates=c, features=f)
    logits = model(field)
    loss = ops.cross_entropy(logits, y)
    tr.backward_and_step(loss)
    return field.coordinate_manager.size(field.coordinate_manager.get_unique_coordinate_map_key(1))


for _ in range(5):
    vox = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"ResNet14 fwd+bwd+SGD  B={args.batch}  {vox} voxels/step  {args.precision}: {ms:.3f} ms/step  "
      f"{vox / ms / 1e3:.2f} M voxels/s  {args.batch / ms * 1e3:.0f} objects/s", flush=True)
if args.cpu:
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
