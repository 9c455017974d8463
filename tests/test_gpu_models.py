"""Whole-network parity: the from-scratch ResNet14 / Res16UNet34C (same parameters and layer tables
as the reference's resnet.py / res16unet.py, see tests/test_dropin.py) on the CUDA engine against the
functional CPU oracle (oracle/nets.py, fp64) with identical weights.

Stated end-to-end bars (14 / 50+ layers with batch-norm in between; measured values in brackets are
ResNet14 / Res16UNet34C on the small random-weight test scenes):
  fp32 : logits |d| <= 2e-3 max|ref|, cos >= 0.99999 ; gradients, all parameters concatenated: cos >= 0.9999
         [1.00000 / 0.99997], worst single parameter >= 0.999
  tf32 : logits cos >= 0.9999, |d| <= 3e-2 max|ref| ; gradients cos >= 0.95 [0.99909 / 0.968],
         worst single parameter >= 0.85 [0.914 .. 0.920 over runs: wgrad sums with fp32 atomics]
  bf16 : logits cos >= 0.995 [0.99999 / 0.99888], |d| <= 2e-1 max|ref| ; gradients cos >= 0.7 [0.993 / 0.80-0.81],
         worst single parameter >= 0.45 [0.57 .. 0.59]
Every individual op inside these backward passes agrees with an fp64 recomputation to its op-level bound
(tests/test_gpu_parity.py: fp32 1e-4, tf32 3e-3, bf16 2e-2 of max|ref|; scripts/diag_ops_in_model.py).  The
looser end-to-end gradient bars measure how operand rounding is amplified through 50 layers of batch-norm
over a few dozen rows at the deepest level of a SMALL scene with random weights and random labels: plain
fp32 rounding (6e-8) already shows up as 3e-5 in the UNet gradient, the same x500 amplification takes
TF32 (5e-4) and bf16 (4e-3) operand rounding to the values above.  They are properties of this test
problem, not of the kernels.
"""
import numpy as np
import pytest
import torch

from nerf_downstream_b200 import models, ops, synth
from nerf_downstream_b200 import me as ME
from oracle import nets

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _compare(model, fwd, coords, feats, target_fn, mode, dev):
    ops.set_default_precision(mode)
    try:
        model = model.to(dev).train()
        params = {k: v.detach().double().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
                  for k, v in model.state_dict().items()}
        ref = fwd(params, coords, torch.from_numpy(feats).double())
        loss_ref = target_fn(ref)
        loss_ref.backward()
        field = ME.TensorField(coordinates=torch.from_numpy(coords).to(dev), features=torch.from_numpy(feats).to(dev))
        out = model(field)
        loss = target_fn(out)
        loss.backward()
        scale = ref.abs().max().item()
        err = (out.detach().double().cpu() - ref.detach()).abs().max().item()
        cos = _cos(out.detach(), ref.detach())
        print(f"[{mode}] logits cos={cos:.6f} max err={err:.3e} (scale {scale:.3e})")
        if mode == "fp32":
            assert err <= 2e-3 * scale, (err, scale)
            assert cos >= 0.99999
        elif mode == "tf32":
            assert cos >= 0.9999, cos
            assert err <= 3e-2 * scale, (err, scale)
        else:
            assert cos >= 0.995, cos
            assert err <= 2e-1 * scale, (err, scale)
        worst, worst_name = 1.0, ""
        ga, gr = [], []
        for name, p in model.named_parameters():
            g_ref = params[name].grad
            assert p.grad is not None and g_ref is not None, name
            c = _cos(p.grad, g_ref)
            if c < worst:
                worst, worst_name = c, name
            ga.append(p.grad.detach().double().flatten().cpu())
            gr.append(g_ref.detach().double().flatten())
        total = _cos(torch.cat(ga), torch.cat(gr))
        print(f"[{mode}] logits cos={cos:.6f} max err={err:.3e} (scale {scale:.3e}) all-parameter grad cos={total:.5f} "
              f"worst single parameter {worst_name}: {worst:.4f}")
        assert total >= {"fp32": 0.9999, "tf32": 0.95, "bf16": 0.7}[mode], total
        assert worst >= {"fp32": 0.999, "tf32": 0.85, "bf16": 0.45}[mode], (worst_name, worst)
        return cos, worst
    finally:
        ops.set_default_precision("tf32")


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_resnet14_matches_oracle(cuda_device, mode):
    torch.manual_seed(0)
    coords, feats, labels = synth.co3d_batch(777, 3, lattice=64)
    model = models.ResNet14(27, 51)
    y = torch.from_numpy(labels)

    def target(logits):
        return torch.nn.functional.cross_entropy(logits, y.to(logits.device))
    _compare(model, nets.resnet_forward, coords, feats, target, mode, cuda_device)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_res16unet34c_matches_oracle(cuda_device, mode):
    torch.manual_seed(1)
    coords, feats, labels = synth.room_batch(777, 2, 40_000)
    model = models.Res16UNet34C(27, 20)
    y = torch.from_numpy(labels)

    def target(logits):
        return torch.nn.functional.cross_entropy(logits, y.to(logits.device), ignore_index=255)
    _compare(model, nets.resunet_forward, coords, feats, target, mode, cuda_device)


def test_manager_semantics(cuda_device):
    """Key rules the models silently rely on (SURVEY.md §7 hard parts)."""
    coords, feats = synth.random_cloud(3, 4000, extent=10, n_batch=2, channels=8)
    x = ME.TensorField(coordinates=torch.from_numpy(coords).to(cuda_device),
                       features=torch.from_numpy(feats).to(cuda_device)).sparse()
    mgr = x.coordinate_manager
    a = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=2, dimension=3).to(cuda_device)(x)
    b = ME.MinkowskiConvolution(8, 8, kernel_size=1, stride=2, dimension=3).to(cuda_device)(x)
    assert a.coordinate_map_key == b.coordinate_map_key and a.tensor_stride == [2, 2, 2]
    a += b                                                      # legal because the keys are equal
    up = ME.MinkowskiConvolutionTranspose(8, 4, kernel_size=2, stride=2, dimension=3).to(cuda_device)(a)
    assert up.coordinate_map_key == x.coordinate_map_key        # lands on the encoder's map
    cat = ME.cat(up, x)
    assert cat.F.shape == (x.F.shape[0], 12)
    with pytest.raises(AssertionError):
        ME.cat(a, x)
    g = ME.MinkowskiGlobalAvgPooling()(a)
    assert g.F.shape == (2, 8) and g.C.cpu().tolist() == [[0, 0, 0, 0], [1, 0, 0, 0]]
    km = mgr.kernel_map(x.coordinate_map_key, x.coordinate_map_key, 1, 3, 1)
    assert 13 in km and km[13].shape[0] == 2 and bool((km[13][0] == km[13][1]).all())
    assert mgr.size(x.coordinate_map_key) == x.F.shape[0]
    k2 = mgr.stride(x.coordinate_map_key, [2, 2, 2])
    assert k2 == a.coordinate_map_key


@pytest.mark.gpu
def test_engine_side_row_order_on_the_faithful_geometry(cuda_device):
    """ops.reorder_rows_by_mask on the real library: a faithful ScanNet-plenoxel scene (few
    neighbours per voxel, raster tiles touching most offsets) is re-ordered at the levels where it pays; logits at the
    points, loss and parameter gradients equal those of the first-occurrence order (fp32 CUDA-core path: only the
    order of sums differs), the re-ordered maps hold the same voxels, and the executed (tile, offset) volume drops."""
    from nerf_downstream_b200 import lib as L
    c, f, y = synth.faithful_room_batch(5, 1, 150_000, scene_scale=0.56)
    c_d, f_d, y_d = (torch.from_numpy(a).to(cuda_device) for a in (c, f, y))
    saved = (ops.sort_rows, ops.sort_min_rows, ops.sort_window, dict(ops.sort_stats))

    def run(sort):
        ops.sort_rows, ops.sort_min_rows, ops.sort_window = sort, 20_000, 16_384
        ops.sort_stats.update(considered=0, reordered=0)
        ops.set_default_precision("fp32")
        torch.manual_seed(8)
        net = models.Res16UNet14A(27, 20).to(cuda_device).train()
        field = ME.TensorField(coordinates=c_d, features=f_d)
        out = net(field)
        loss = ops.cross_entropy(out, y_d, 255)
        loss.backward()
        torch.cuda.synchronize()
        mgr = field.coordinate_manager
        vol = {}
        for ts in (1, 2):
            key = mgr.get_unique_coordinate_map_key(ts)
            km = mgr.get_kernel_map(key, key, ME.KernelGenerator(kernel_size=3, stride=1, dilation=1, dimension=3))
            vol[ts] = (int(sum(bin(int(v) & 0xFFFFFFFF).count("1") for v in km.mask.tolist())) * 128, km.n_pairs,
                       mgr.get_coordinates(key).clone())
        return out.detach().clone(), float(loss), torch.cat([p.grad.flatten() for p in net.parameters()]), vol, dict(ops.sort_stats)

    try:
        out_a, loss_a, g_a, vol_a, st_a = run(False)
        out_b, loss_b, g_b, vol_b, st_b = run(True)
    finally:
        ops.sort_rows, ops.sort_min_rows, ops.sort_window = saved[:3]
        ops.sort_stats.update(saved[3])
        ops.set_default_precision("tf32")
    assert st_a["reordered"] == 0 and st_b["reordered"] >= 1, (st_a, st_b)
    assert (out_a - out_b).abs().max() <= 1e-4 * out_a.abs().max()
    assert abs(loss_a - loss_b) <= 1e-5 * abs(loss_a)
    a, b = g_a.double(), g_b.double()
    assert float(a @ b / (a.norm() * b.norm())) >= 0.99999
    improved = 0
    for ts in (1, 2):
        ex_a, p_a, ca = vol_a[ts]
        ex_b, p_b, cb = vol_b[ts]
        assert p_a == p_b and ca.shape == cb.shape
        ka = (ca[:, 0].long() << 54) | ((ca[:, 1].long() + 131072) << 36) | ((ca[:, 2].long() + 131072) << 18) | (ca[:, 3].long() + 131072)
        kb = (cb[:, 0].long() << 54) | ((cb[:, 1].long() + 131072) << 36) | ((cb[:, 2].long() + 131072) << 18) | (cb[:, 3].long() + 131072)
        assert torch.equal(torch.sort(ka)[0], torch.sort(kb)[0])          # the same voxels
        assert ex_b <= ex_a
        improved += ex_b < 0.8 * ex_a
        print(f"ts{ts}: executed / useful {ex_a / p_a:.2f} -> {ex_b / p_b:.2f}")
    assert improved >= 1
