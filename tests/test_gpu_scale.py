"""Operand precision at REALISTIC scale: Res16UNet34C on 2 x 200 K-voxel ScanNet-shaped scenes,
the fp32 CUDA-core path (SPC_PREC_FP32, FFMA) as the on-device reference.

Per layer ("TF32/bf16 tensor-core mode within 3e-3 * max|ref| per layer"): conv fwd / dgrad /
wgrad of the UNet's layer shapes on the scene's own 3^3 map — bf16 measured 2.2e-3 .. 2.7e-3, tf32 7.2e-4 .. 8.5e-4.

Whole network (same weights, same batch, loss = cross-entropy with 10 % ignored labels): measured on a B200
(profiles/r2_precision_at_scale.md)
    tf32: logits cos 0.999986, max|d| 7.3e-3 max|ref|, all-parameter gradient cos 0.985, worst single parameter 0.938
    bf16: logits cos 0.99933,  max|d| 5.0e-2 max|ref|, all-parameter gradient cos 0.907, worst single parameter 0.665
    fp32 run against itself: identical (cos 1.0) — the reference is deterministic.
The gradient of this random-weight, random-label problem amplifies operand rounding ~x500 (fp32's own 6e-8 shows up
as 3e-5, tests/test_gpu_models.py), so even the TF32 mode BASELINE.md blesses per layer does not reach a 0.99
all-parameter cosine; the bars below are the measured values with margin.  What the bars guard is that the
tensor-core paths stay AT those values — a kernel bug (wrong tile, lost offset) drops the cosine to ~0.
"""
import pytest
import torch

from nerf_downstream_b200 import lib as L
from nerf_downstream_b200 import me as ME
from nerf_downstream_b200 import models, ops, synth

pytestmark = pytest.mark.gpu

VOXELS, SCENES = 200_000, 2
PER_LAYER_BAR = 3e-3


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


@pytest.fixture(scope="module")
def scene(cuda_device):
    coords, feats, labels = synth.room_batch(777, SCENES, VOXELS)
    return tuple(torch.from_numpy(a).to(cuda_device) for a in (coords, feats, labels))


@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 96), (128, 96), (64, 64), (256, 256)])
def test_per_layer_bound_on_a_400k_row_map(cuda_device, scene, cin, cout):
    c_d = scene[0]
    cmap, _, _, _ = ops.coords_insert(c_d, L.SRC_FLOAT, (1, 1, 1))
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    assert cmap.size >= 390_000
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout)
    x = torch.randn(cmap.size, cin, generator=g).to(cuda_device)
    w = (torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5).to(cuda_device)
    go = torch.randn(cmap.size, cout, generator=g).to(cuda_device)
    ref = (ops.conv_fwd_raw(x, w, None, km, L.PREC_FP32), ops.conv_dgrad_raw(go, w, km, L.PREC_FP32),
           ops.conv_wgrad_raw(x, go, km, 27, cin, cout, L.PREC_FP32))
    for prec, name in ((L.PREC_BF16, "bf16"), (L.PREC_TF32, "tf32")):
        xa, ga = (ops.to_bf16(x), ops.to_bf16(go)) if prec == L.PREC_BF16 else (x, go)
        got = (ops.conv_fwd_raw(xa, w, None, km, prec), ops.conv_dgrad_raw(ga, w, km, prec),
               ops.conv_wgrad_raw(xa, ga, km, 27, cin, cout, prec))
        for what, a, b in zip(("fwd", "dgrad", "wgrad"), got, ref):
            err = float((a - b).abs().max() / b.abs().max())
            assert err <= PER_LAYER_BAR, f"{name} {what} {cin}->{cout}: |d| = {err:.2e} max|ref| > {PER_LAYER_BAR}"


def _run(mode, scene, dev):
    c_d, f_d, y_d = scene
    ops.set_default_precision(mode)
    try:
        torch.manual_seed(1)
        model = models.Res16UNet34C(27, 20).to(dev).train()
        out = model(ME.TensorField(coordinates=c_d, features=f_d))
        loss = torch.nn.functional.cross_entropy(out, y_d, ignore_index=255)
        loss.backward()
        torch.cuda.synchronize()
        return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    finally:
        ops.set_default_precision("tf32")


# (logits cos, logits max|d| / max|ref|, all-parameter gradient cos, worst single parameter cos)
WHOLE_NET_BARS = {"tf32": (0.9999, 3e-2, 0.97, 0.85), "bf16": (0.998, 1.5e-1, 0.85, 0.5)}


def test_res16unet34c_at_scale_against_the_fp32_path(cuda_device, scene):
    ref_out, ref_g = _run("fp32", scene, cuda_device)
    again_out, again_g = _run("fp32", scene, cuda_device)
    assert _cos(again_out, ref_out) >= 0.999999          # the on-device reference reproduces itself
    for mode, (b_cos, b_err, b_gcos, b_worst) in WHOLE_NET_BARS.items():
        out, grads = _run(mode, scene, cuda_device)
        lcos = _cos(out, ref_out)
        lerr = float((out - ref_out).abs().max() / ref_out.abs().max())
        per = {n: _cos(grads[n], ref_g[n]) for n in ref_g}
        worst = min(per, key=per.get)
        total = _cos(torch.cat([grads[n].flatten() for n in ref_g]), torch.cat([ref_g[n].flatten() for n in ref_g]))
        print(f"[{mode}] logits cos {lcos:.6f} max|d|/max|ref| {lerr:.2e} all-parameter grad cos {total:.4f} "
              f"worst {worst} {per[worst]:.4f}")
        assert lcos >= b_cos and lerr <= b_err, (mode, lcos, lerr)
        assert total >= b_gcos, (mode, total)
        assert per[worst] >= b_worst, (mode, worst, per[worst])
