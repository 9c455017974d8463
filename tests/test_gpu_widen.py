"""GPU parity of the widened rows, through the C ABI, against the CPU oracle:
the fused segmentation head (slice -> SegLoss -> IoUMeter in one pass), class-weighted cross-entropy, instance
normalisation, `ME.utils.sparse_quantize`, and the gin-driven training loop on the real kernels.

Bars: integer outputs (counts, voxel coordinates, index maps, merged labels) exact; loss `<= 1e-5` relative;
gradients / normalised features `|d| <= 1e-4 * (1 + |ref|)` against the fp64 oracle.
"""
import numpy as np
import pytest
import torch

from nerf_downstream_b200 import ginlite, models, ops, pipeline, synth, training
from nerf_downstream_b200 import me as ME
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def close(got, ref, tol=1e-4):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs()
    assert bool((err <= tol * (1 + ref.abs())).all()), f"max err {err.max().item():.3e}"


# ---- f3: segmentation head ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,C,weighted,with_counts", [(1, 1, 20, False, True), (700, 2000, 20, False, True),
                                                        (5000, 5000, 20, True, True), (30_000, 100_000, 20, True, False),
                                                        (900, 4097, 51, True, True), (64, 300, 3, False, False)])
def test_seg_head_matches_reference_steps(cuda_device, m, n, C, weighted, with_counts):
    rng = np.random.default_rng(m + n + C)
    logits = rng.standard_normal((m, C)).astype(np.float32) * 3
    inverse = rng.integers(0, m, n).astype(np.int32)
    inverse[: min(m, n)] = rng.permutation(m)[: min(m, n)]
    target = rng.integers(0, C, n)
    target[rng.random(n) < 0.1] = -255
    target[0] = C - 1                                             # at least one valid point (else the loss is NaN)
    w = None
    if weighted:
        w = torch.ones(C, dtype=torch.float64)
        w[-1] = 0.25
        w[0] = 2.0
    ref_loss, ref_grad, ref_counts = R.seg_head(torch.from_numpy(logits).double(), inverse, target, -255, w)
    x = torch.from_numpy(logits).to(cuda_device).requires_grad_(True)
    counts = torch.zeros((3, C), dtype=torch.int64, device=cuda_device) if with_counts else None
    loss = ops.seg_head(x, torch.from_numpy(inverse).to(cuda_device), torch.from_numpy(target).to(cuda_device), -255,
                        None if w is None else w.float().to(cuda_device), counts)
    (3.0 * loss).backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    close(x.grad, 3.0 * ref_grad)
    if with_counts:
        assert (counts.cpu().numpy() == ref_counts).all()
        # counts accumulate across calls (IoUMeter.update semantics)
        ops.seg_head(x.detach(), torch.from_numpy(inverse).to(cuda_device), torch.from_numpy(target).to(cuda_device),
                     -255, None, counts)
        assert (counts.cpu().numpy() == 2 * ref_counts).all()


@pytest.mark.parametrize("n,C", [(1, 20), (4097, 20), (60_000, 21), (300, 51)])
def test_weighted_cross_entropy_matches_torch(cuda_device, n, C):
    g = torch.Generator().manual_seed(n)
    logits = torch.randn(n, C, generator=g) * 2
    target = torch.randint(0, C, (n,), generator=g)
    target[::9] = -255
    if n > 1:
        target[1] = C - 1
    w = torch.ones(C)
    w[-1] = 0.1                                                     # SegLoss(void_weight=0.1)
    a = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(a, target, weight=w.double(), ignore_index=-255)
    ref.backward()
    b = logits.to(cuda_device).requires_grad_(True)
    crit = training.SegLoss(-255, C, void_weight=0.1).to(cuda_device)
    loss = crit(b, {"labels": target.to(cuda_device)})
    loss.backward()
    if torch.isnan(ref):
        assert torch.isnan(loss)
        return
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    close(b.grad, a.grad)
    # unit weights go through spc_ce_fwd and agree with the weighted kernel run with ones
    plain = training.SegLoss(-255, C).to(cuda_device)(b.detach(), {"labels": target.to(cuda_device)})
    ones = ops.cross_entropy(b.detach(), target.to(cuda_device), -255, torch.ones(C, device=cuda_device))
    assert abs(plain.item() - ones.item()) <= 1e-6 * abs(plain.item())


def test_seg_head_flags_bad_inputs(cuda_device):
    logits = torch.randn(10, 5, device=cuda_device)
    inv = torch.arange(10, dtype=torch.int32, device=cuda_device)
    tgt = torch.zeros(10, dtype=torch.int64, device=cuda_device)
    with pytest.raises(RuntimeError, match="int64"):
        ops.seg_head(logits, inv, tgt.int(), -1)
    with pytest.raises(RuntimeError, match="int32"):
        ops.seg_head(logits, inv.long(), tgt, -1)
    with pytest.raises(RuntimeError, match="1..64"):
        ops.seg_head(torch.randn(4, 70, device=cuda_device), inv[:4], tgt[:4], -1)
    # all points ignored -> NaN (as torch), zero gradient
    x = logits.clone().requires_grad_(True)
    loss = ops.seg_head(x, inv, torch.full((10,), -1, dtype=torch.int64, device=cuda_device), -1)
    assert torch.isnan(loss)


def test_unet_fused_head_equals_slice_then_loss(cuda_device):
    """Res16UNet14A on a small scene: forward_sparse + seg_head_loss == forward (slice) + SegLoss + IoUMeter, for the
    loss, every parameter gradient and the counts."""
    ops.set_default_precision("fp32")
    try:
        torch.manual_seed(0)
        coords, feats, labels = synth.room_batch(5, 1, 6000, ignore_label=-255)
        net = models.Res16UNet14A(27, 20).to(cuda_device).train()
        c, f = torch.from_numpy(coords).to(cuda_device), torch.from_numpy(feats).to(cuda_device)
        y = torch.from_numpy(labels).long().to(cuda_device)
        w = torch.ones(20, device=cuda_device)
        w[-1] = 0.3

        field = ME.TensorField(coordinates=c, features=f)
        logits = net(field)
        loss_a = ops.cross_entropy(logits, y, -255, w)
        meter_a = pipeline.IoUMeter(20, -255)
        meter_a.update(logits.detach(), y)
        net.zero_grad()
        loss_a.backward()
        grads_a = [p.grad.clone() for p in net.parameters()]

        field = ME.TensorField(coordinates=c, features=f)
        out = net.forward_sparse(field)
        meter_b = pipeline.IoUMeter(20, -255)
        loss_b = pipeline.seg_head_loss(out, field, y, -255, w, meter_b.counts_buffer(cuda_device))
        net.zero_grad()
        loss_b.backward()
        assert abs(loss_a.item() - loss_b.item()) <= 1e-5 * abs(loss_a.item())
        assert torch.equal(meter_a.counts, meter_b.counts)
        for ga, p in zip(grads_a, net.parameters()):
            scale = ga.abs().max().item() + 1e-12
            assert (ga - p.grad).abs().max().item() <= 2e-3 * scale
    finally:
        ops.set_default_precision("tf32")


# ---- f4: instance norm, sparse_quantize --------------------------------------------------------------------------
@pytest.mark.parametrize("m,C,n_batch", [(1, 4, 1), (1000, 32, 3), (5003, 96, 4), (20_000, 27, 2), (300, 130, 5),
                                         (3000, 64, 3), (600, 1028, 2), (100_000, 128, 2)])
def test_instance_norm_matches_oracle(cuda_device, m, C, n_batch):
    rng = np.random.default_rng(m + C)
    batch = np.sort(rng.integers(0, n_batch, m)).astype(np.int32)
    if m > 100:                                                   # rows of an instance need not be contiguous
        batch[m // 2: m // 2 + 7] = 0
    coords = np.zeros((m, 4), np.int32)
    coords[:, 0] = batch
    coords[:, 1:] = rng.integers(-50, 50, (m, 3))
    x = (rng.standard_normal((m, C)) * 2 + 0.7).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal((1, C))).astype(np.float32)
    beta = (0.1 * rng.standard_normal((1, C))).astype(np.float32)
    gy = rng.standard_normal((m, C)).astype(np.float32)

    xr = torch.from_numpy(x).double().requires_grad_(True)
    gr = torch.from_numpy(gamma).double().requires_grad_(True)
    br = torch.from_numpy(beta).double().requires_grad_(True)
    ref = R.instance_norm(xr, batch, n_batch, gr, br, eps=1e-8)
    ref.backward(torch.from_numpy(gy).double())

    xg = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
    gg = torch.from_numpy(gamma).to(cuda_device).requires_grad_(True)
    bg = torch.from_numpy(beta).to(cuda_device).requires_grad_(True)
    out = ops.InstanceNormFn.apply(xg, torch.from_numpy(coords).to(cuda_device), n_batch, gg, bg, 1e-8)
    out.backward(torch.from_numpy(gy).to(cuda_device))
    if m == 1:                                                    # one row: var 0 -> rstd 1e4, output = beta
        close(out, ref, 1e-3)
        return
    close(out, ref)
    close(xg.grad, xr.grad, 2e-4)
    close(gg.grad, gr.grad, 2e-4)
    close(bg.grad, br.grad, 2e-4)


def test_instance_norm_vector_and_scalar_kernels_agree(cuda_device):
    from nerf_downstream_b200 import lib as L
    rng = np.random.default_rng(9)
    m, C, nb = 7001, 96, 3
    coords = np.zeros((m, 4), np.int32)
    coords[:, 0] = np.sort(rng.integers(0, nb, m))
    x = torch.from_numpy(rng.standard_normal((m, C)).astype(np.float32)).to(cuda_device)
    gy = torch.from_numpy(rng.standard_normal((m, C)).astype(np.float32)).to(cuda_device)
    cg = torch.from_numpy(coords).to(cuda_device)
    res = []
    try:
        for scalar in (0, 1):
            L.load().spc_inst_norm_force_scalar(scalar)
            xg = x.clone().requires_grad_(True)
            gam = torch.full((1, C), 1.5, device=cuda_device, requires_grad=True)
            bet = torch.full((1, C), -0.5, device=cuda_device, requires_grad=True)
            out = ops.InstanceNormFn.apply(xg, cg, nb, gam, bet, 1e-8)
            out.backward(gy)
            res.append((out.detach(), xg.grad, gam.grad, bet.grad))
    finally:
        L.load().spc_inst_norm_force_scalar(0)
    for a, b in zip(*res):
        close(a, b, 2e-5)


def test_instance_norm_module_on_sparse_tensor(cuda_device):
    coords, feats, _ = synth.co3d_batch(3, 3, lattice=24)
    field = ME.TensorField(coordinates=torch.from_numpy(coords).to(cuda_device),
                           features=torch.from_numpy(feats).to(cuda_device))
    s = field.sparse()
    norm = ME.MinkowskiInstanceNorm(27).to(cuda_device)
    assert set(norm.state_dict()) == {"weight", "bias"} and tuple(norm.weight.shape) == (1, 27)
    out = norm(s)
    assert out.coordinate_map_key == s.coordinate_map_key
    batch = s.C[:, 0].cpu().numpy()
    ref = R.instance_norm(s.F.detach().double().cpu(), batch, 3)
    close(out.F, ref)
    for b in range(3):                                            # zero mean, unit variance per instance and channel
        rows = out.F[s.C[:, 0] == b]
        assert rows.mean(0).abs().max().item() < 1e-4 and (rows.var(0, unbiased=False) - 1).abs().max().item() < 1e-3


@pytest.mark.parametrize("n,q,dtype", [(1, 0.5, np.float64), (5000, 0.05, np.float64), (20_000, 0.02, np.float32),
                                       (3000, None, np.float32)])
def test_sparse_quantize_exact(cuda_device, n, q, dtype):
    rng = np.random.default_rng(n)
    xyz = rng.uniform(-1.5, 1.5, (n, 3)).astype(dtype) * (1 if q else 20)
    colors = rng.uniform(0, 255, (n, 3)).astype(np.float32)
    labels = rng.integers(0, 4, n).astype(np.int32)
    rc, rf, rl, rfirst, rinv = R.sparse_quantize_np(xyz, colors, labels, -100, q)
    c, f, lab, idx, inv = ME.utils.sparse_quantize(xyz, colors, labels=labels, quantization_size=q, return_index=True,
                                                   return_inverse=True, ignore_label=-100)
    assert isinstance(c, np.ndarray) and c.dtype == np.int32               # numpy in -> numpy out, as ME
    assert (c == rc).all() and (idx == rfirst).all() and (inv == rinv).all()
    assert (f == rf).all() and (lab == rl).all()
    if n > 1000 and q:
        assert (rl == -100).any()                                          # label conflicts were exercised
    # the call of scannet.py:235-242 and the maps-only form
    _, f2, l2, rows = ME.utils.sparse_quantize(np.ascontiguousarray(xyz), colors, labels=labels, quantization_size=q,
                                               return_index=True, ignore_label=-100)
    assert (rows == rfirst).all() and (l2 == rl).all()
    um = ME.utils.sparse_quantize(torch.from_numpy(xyz).to(cuda_device), return_maps_only=True, quantization_size=q)
    assert torch.is_tensor(um) and (um.cpu().numpy() == rfirst).all()


# ---- f1: the gin-driven loop on the real kernels -------------------------------------------------------------------
@pytest.mark.parametrize("fused", [False, True])
def test_gin_driven_segmentation_run(cuda_device, tmp_path, fused):
    ginlite.clear_config()
    try:
        cfg_file = tmp_path / "tiny.gin"
        cfg_file.write_text("""
get_model.name = "Res16UNet14A"
get_model.in_channel = 27
get_model.out_channel = 20
train.max_steps = 6
train.warmup_steps = 2
train.scheduler_name = "PolyLR"
PolyLR.poly_exp = 0.9
train.lr = 0.05
train.weight_decay = 1e-4
train.ignore_label = -255
train.val_every_n_steps = 4
train.log_every_n_steps = 2
train.void_weight = 0.5
SGD.momentum = 0.9
""")
        torch.manual_seed(0)

        def batches(seed, k):
            out = []
            for i in range(k):
                c, f, y = synth.room_batch(seed + i, 1, 4000, ignore_label=-255)
                out.append({"coordinates": torch.from_numpy(c).to(cuda_device), "features": torch.from_numpy(f).to(cuda_device),
                            "labels": torch.from_numpy(y).to(cuda_device)})
            return out
        train_b, val_b = batches(100, 3), batches(200, 2)
        logs = []
        run = training.train([str(cfg_file)], ["train.gpus=1"], lambda: train_b, lambda: val_b,
                             save_path=str(tmp_path / "run"), device=cuda_device, log=logs.append, fused_head=fused)
        assert run.global_step == 8 and isinstance(run.model, models.Res16UNet14A)
        sched = run.cfg.schedule()
        assert abs(run.trainer.lr - sched.lr(8)) < 1e-12
        tl = [d for d in logs if "train/loss" in d]
        assert [d["global_step"] for d in tl] == [2, 4, 6] and all(np.isfinite(d["train/loss"]) for d in tl)
        vl = [d for d in logs if "val/mIoU" in d]
        assert [d["global_step"] for d in vl] == [4, 8] and 0 <= vl[-1]["val/mIoU"] <= 100
        ckpt = torch.load(tmp_path / "run" / "last.ckpt", weights_only=False)
        assert ckpt["global_step"] == 8 and all(k.startswith("model.") for k in ckpt["state_dict"])
        # the checkpoint evaluates in a fresh model exactly like the trained one (eval.py:47-67)
        fresh = models.Res16UNet14A(27, 20).to(cuda_device).eval()
        training.load_lightning_state_dict(fresh, ckpt["state_dict"])
        run.model.eval()
        with torch.no_grad():
            b = val_b[0]
            a = run.model(ME.TensorField(coordinates=b["coordinates"], features=b["features"]))
            z = fresh(ME.TensorField(coordinates=b["coordinates"], features=b["features"]))
        # (small maps split the kernel offsets over CTAs and meet in fp32 atomics: equal up to summation order)
        assert (a - z).abs().max().item() <= 1e-3 * a.abs().max().item()
    finally:
        ginlite.clear_config()


# ---- f4: trilinear interpolation / splat ---------------------------------------------------------------------------
@pytest.mark.parametrize("n_vox,n_query,C,ts", [(3000, 5000, 8, 1), (3000, 4000, 3, 2), (200, 1, 5, 1),
                                                 (20_000, 60_000, 32, 1)])
def test_interpolation_matches_oracle(cuda_device, n_vox, n_query, C, ts):
    rng = np.random.default_rng(n_vox + n_query)
    c, _ = synth.random_cloud(n_vox, n_vox, extent=10, n_batch=2)
    cmap, _, _, _ = ops.coords_insert(torch.from_numpy(c).to(cuda_device), 0, (ts, ts, ts))       # 0 = SRC_FLOAT
    map_coords = cmap.coords.cpu().numpy()
    q = np.empty((n_query, 4), np.float32)
    q[:, 0] = rng.integers(0, 2, n_query)
    q[:, 1:] = rng.uniform(-10.5, 10.5, (n_query, 3))
    q[: n_query // 10, 1:] = np.round(q[: n_query // 10, 1:])                   # points exactly on the lattice
    feats = rng.standard_normal((cmap.size, C)).astype(np.float32)
    gy = rng.standard_normal((n_query, C)).astype(np.float32)

    _, rows, w = R.interp_map_np(map_coords, q, (ts, ts, ts))
    idx_g, w_g = ops.interp_map(cmap, torch.from_numpy(q).to(cuda_device))
    assert (idx_g.cpu().numpy() == rows).all()                                   # integer part exact
    assert np.abs(w_g.cpu().numpy() - w).max() <= 1e-6
    fr = torch.from_numpy(feats).double().requires_grad_(True)
    ref = R.interpolate(fr, rows, w)
    ref.backward(torch.from_numpy(gy).double())
    fg = torch.from_numpy(feats).to(cuda_device).requires_grad_(True)
    out = ops.InterpolateFn.apply(fg, idx_g, w_g)
    out.backward(torch.from_numpy(gy).to(cuda_device))
    close(out, ref)
    close(fg.grad, fr.grad)


def test_splat_and_interpolate_on_me_surface(cuda_device):
    rng = np.random.default_rng(3)
    n, C = 4000, 6
    q = np.empty((n, 4), np.float32)
    q[:, 0] = np.sort(rng.integers(0, 2, n))
    q[:, 1:] = rng.uniform(-6, 6, (n, 3))
    f = rng.standard_normal((n, C)).astype(np.float32)
    x = ME.TensorField(coordinates=torch.from_numpy(q).to(cuda_device),
                       features=torch.from_numpy(f).to(cuda_device).requires_grad_(True))
    y = x.splat()
    uc, ref, rows, w = R.splat(torch.from_numpy(f).double(), q)
    assert (y.C.cpu().numpy() == uc).all() and y.tensor_stride == [1, 1, 1]      # voxels in order of first touch
    close(y.F, ref)
    assert abs(y.F.sum().item() - f.sum()) <= 1e-2                               # weights of a point sum to one
    # back to the points: sum_k w_k * splat[row_k]
    back = y.interpolate(x)
    assert back.coordinate_field_map_key == x.coordinate_field_map_key
    close(back.F, R.interpolate(ref, rows, w), 2e-4)
    back.F.sum().backward()                                                      # gradient flows through both maps
    wsum = torch.from_numpy(w).double().sum(0)
    assert x.F.grad is not None and x.F.grad.shape == (n, C)
    # d/df_j of sum_i back_i = sum over voxels v touched by j of w_jv * (sum_i w_iv)
    colsum = torch.zeros(uc.shape[0], dtype=torch.float64)
    for k in range(8):
        colsum.index_add_(0, torch.from_numpy(rows[k]).long(), torch.from_numpy(w[k]).double())
    want = sum(torch.from_numpy(w[k]).double() * colsum[torch.from_numpy(rows[k]).long()] for k in range(8))
    close(x.F.grad[:, 0], want, 2e-4)
    # MinkowskiInterpolation on a strided tensor (a linear function on a full lattice is reproduced exactly)
    g = np.stack(np.meshgrid(*[np.arange(-4, 5, 2)] * 3, indexing="ij"), -1).reshape(-1, 3)
    bc = np.concatenate([np.zeros((len(g), 1)), g], 1).astype(np.int32)
    lin = (g @ np.array([[1.0], [-2.0], [0.5]]) + 3.0).astype(np.float32)
    s = ME.SparseTensor(torch.from_numpy(lin).to(cuda_device), coordinates=torch.from_numpy(bc).to(cuda_device),
                        tensor_stride=2)
    pts = np.concatenate([np.zeros((100, 1)), rng.uniform(-4, 3.9, (100, 3))], 1).astype(np.float32)
    out, (in_rows, out_rows), wts = ME.MinkowskiInterpolation(return_kernel_map=True, return_weights=True)(
        s, torch.from_numpy(pts).to(cuda_device))
    want = pts[:, 1:] @ np.array([[1.0], [-2.0], [0.5]]) + 3.0
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-4
    assert in_rows.shape == out_rows.shape == wts.shape and in_rows.numel() == 800


# ---- f2: decode + the reference's affine augmentations in one pass -----------------------------------------------
@pytest.mark.parametrize("seed,names", [(0, ["RandomRotation", "RandomAffine", "RandomHorizontalFlip", "RandomTranslation"]),
                                        (3, ["RandomRotation", "RandomScale", "CoordinateUniformTranslation"]),
                                        (5, [])])
def test_plenoxel_decode_augmented(cuda_device, seed, names):
    """The composed chain runs inside spc_plenoxel_decode: bit-exact against the numpy decode given the same 12 floats,
    and equal (to float32 rounding) to the float64 chain applied to the decoded lattice indices."""
    import random
    rng = np.random.default_rng(seed)
    reso = (128, 128, 128)
    n = 20_000
    links = np.sort(rng.choice(reso[0] * reso[1] * reso[2], size=n, replace=False)).astype(np.int64)
    sh = rng.integers(0, 256, size=(n, 27), dtype=np.uint8)
    scale, mn = np.float32(2.0 / 255.0), np.float32(-1.0)
    params = {"RandomRotation": dict(upright_axis="y", application_ratio=1.0),
              "RandomAffine": dict(upright_axis="y", application_ratio=1.0),
              "RandomHorizontalFlip": dict(upright_axis="y", application_ratio=1.0),
              "RandomTranslation": dict(max_translation=3, application_ratio=1.0),
              "RandomScale": dict(scale_ratio=0.4, application_ratio=1.0),
              "CoordinateUniformTranslation": dict(max_translation=0.2)}
    random.seed(seed)
    np.random.seed(seed)
    c, f, chain = pipeline.plenoxel_decode_augmented(torch.from_numpy(links).to(cuda_device),
                                                     torch.from_numpy(sh).to(cuda_device), float(scale), float(mn), reso,
                                                     names, batch_index=1, params=params)
    assert len(chain.steps) >= len(names)
    rc, rf = R.plenoxel_decode_np(links, sh, scale, mn, reso, batch_index=1,
                                  affine=chain.as_affine12() if chain.steps else None)
    assert (c.cpu().numpy() == rc).all() and (f.cpu().numpy() == rf).all()
    ijk = np.stack([links // (128 * 128), (links % (128 * 128)) // 128, links % 128], 1).astype(np.float64)
    want = chain.apply(ijk)
    assert np.abs(c[:, 1:].cpu().numpy() - want).max() <= 2e-4 * max(1.0, np.abs(want).max())
    if "RandomHorizontalFlip" in names:                            # mirrored about the data's own maximum
        assert {"flip0", "flip2"} <= set(chain.steps)
