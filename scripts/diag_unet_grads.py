"""Per-parameter gradient agreement (CUDA fp32 mode vs fp64 oracle) for Res16UNet34C, printed in
backward order, to localise a backward bug."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth  # noqa: E402
from oracle import nets  # noqa: E402

voxels = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
torch.manual_seed(1)
ops.set_default_precision(mode)
coords, feats, labels = synth.room_batch(777, 2, voxels)
model = models.Res16UNet34C(27, 20).to(dev).train()
params = {k: v.detach().double().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in model.state_dict().items()}
y = torch.from_numpy(labels)
ref = nets.resunet_forward(params, coords, torch.from_numpy(feats).double())
torch.nn.functional.cross_entropy(ref, y, ignore_index=255).backward()
field = ME.TensorField(coordinates=torch.from_numpy(coords).to(dev), features=torch.from_numpy(feats).to(dev))
out = model(field)
torch.nn.functional.cross_entropy(out, y.to(dev), ignore_index=255).backward()
mgr = field.coordinate_manager
print("maps:", {str(k): m.size for k, m in mgr._maps.items()})
print("logit err", (out.detach().double().cpu() - ref.detach()).abs().max().item(), "scale", ref.abs().max().item())
for name, p in reversed(list(model.named_parameters())):
    a, b = p.grad.double().cpu().flatten(), params[name].grad.flatten()
    cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
    print(f"{cos:10.6f}  {float(a.norm()):10.3e} {float(b.norm()):10.3e}  {name}")
