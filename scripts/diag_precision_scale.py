"""Operand-precision study at REALISTIC scale: Res16UNet34C on >= 200 K-voxel scenes, the
fp32 CUDA-core path (PREC_FP32, FFMA) as the on-device reference, bf16 / tf32 tensor-core modes against it.

Prints (a) per-layer |d| / max|ref| of conv fwd / dgrad / wgrad on the scene's own 3^3 map for the UNet's layer
shapes, (b) logits cosine, all-parameter gradient cosine and the worst single parameter for every mode.

  python scripts/diag_precision_scale.py [voxels_per_scene] [scenes]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import lib as L  # noqa: E402
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth  # noqa: E402

voxels = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
coords, feats, labels = synth.room_batch(777, scenes, voxels)
c_d, f_d, y_d = (torch.from_numpy(a).to(dev) for a in (coords, feats, labels))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


# ---- (a) per-layer bound on the scene's stride-1 3^3 map ---------------------------------------
cmap, _, _, _ = ops.coords_insert(c_d, L.SRC_FLOAT, (1, 1, 1))
km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
g = torch.Generator(device="cpu").manual_seed(0)
layer_rows = []
for cin, cout in ((32, 32), (96, 96), (128, 96), (64, 64), (256, 256)):
    x = torch.randn(cmap.size, cin, generator=g).to(dev)
    w = (torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5).to(dev)
    go = torch.randn(cmap.size, cout, generator=g).to(dev)
    ref = (ops.conv_fwd_raw(x, w, None, km, L.PREC_FP32), ops.conv_dgrad_raw(go, w, km, L.PREC_FP32),
           ops.conv_wgrad_raw(x, go, km, 27, cin, cout, L.PREC_FP32))
    for prec, name in ((L.PREC_BF16, "bf16"), (L.PREC_TF32, "tf32")):
        if prec == L.PREC_BF16:
            xa, ga = ops.to_bf16(x), ops.to_bf16(go)
        else:
            xa, ga = x, go
        got = (ops.conv_fwd_raw(xa, w, None, km, prec), ops.conv_dgrad_raw(ga, w, km, prec),
               ops.conv_wgrad_raw(xa, ga, km, 27, cin, cout, prec))
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, ref)]
        layer_rows.append({"layer": f"{cin}->{cout}", "mode": name, "rows": cmap.size,
                           "fwd": errs[0], "dgrad": errs[1], "wgrad": errs[2]})
        print(f"layer {cin:3d}->{cout:3d} {name}: |d|/max|ref| fwd {errs[0]:.2e} dgrad {errs[1]:.2e} wgrad {errs[2]:.2e}",
              flush=True)
    del x, w, go, ref


# ---- (b) whole network ------------------------------------------------------------------------
def run(mode):
    ops.set_default_precision(mode)
    torch.manual_seed(1)
    model = models.Res16UNet34C(27, 20).to(dev).train()
    field = ME.TensorField(coordinates=c_d, features=f_d)
    out = model(field)
    loss = torch.nn.functional.cross_entropy(out, y_d, ignore_index=255)
    loss.backward()
    torch.cuda.synchronize()
    return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}, float(loss)


ref_out, ref_g, ref_loss = run("fp32")
summary = {"voxels_per_scene": voxels, "scenes": scenes, "rows_ts1": cmap.size, "layers": layer_rows, "modes": {}}
for mode in ("fp32", "tf32", "bf16"):
    out, grads, loss = run(mode)
    per = {n: cos(grads[n], ref_g[n]) for n in ref_g}
    worst = min(per, key=per.get)
    total = cos(torch.cat([grads[n].flatten() for n in ref_g]), torch.cat([ref_g[n].flatten() for n in ref_g]))
    lerr = float((out - ref_out).abs().max() / ref_out.abs().max())
    conv_only = [n for n in ref_g if n.endswith("kernel")]
    total_conv = cos(torch.cat([grads[n].flatten() for n in conv_only]), torch.cat([ref_g[n].flatten() for n in conv_only]))
    summary["modes"][mode] = {"logits_cos": cos(out, ref_out), "logits_maxerr_rel": lerr, "grad_cos_all": total,
                              "grad_cos_conv_kernels": total_conv, "worst_param": worst, "worst_cos": per[worst],
                              "loss": loss, "ref_loss": ref_loss,
                              "n_params_below_0.99": sum(1 for v in per.values() if v < 0.99)}
    print(f"[{mode}] vs fp32 CUDA-core run: logits cos {cos(out, ref_out):.6f} max|d|/max|ref| {lerr:.2e}  "
          f"all-parameter grad cos {total:.5f}  worst {worst} {per[worst]:.4f}  "
          f"params<0.99: {summary['modes'][mode]['n_params_below_0.99']}/{len(per)}", flush=True)
ops.set_default_precision("tf32")
print("JSON " + json.dumps(summary))
