"""Oracle pinned on hand-checkable known answers and on itself
(numpy restatement vs C restatement).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R


def test_offset_numbering_z_axis():
    # sparse_conv.py:375-379: valid_kernel=[4,13,22] is "the z axis" of a 3x3x3 kernel
    offs = R.kernel_offsets((3, 3, 3), (1, 1, 1))
    assert len(offs) == 27
    assert offs[4] == (0, 0, -1) and offs[13] == (0, 0, 0) and offs[22] == (0, 0, 1)
    assert offs[0] == (-1, -1, -1) and offs[1] == (0, -1, -1) and offs[3] == (-1, 0, -1)
    # even kernels start at 0, offsets scale with the input tensor stride
    assert R.kernel_offsets((2, 2, 2), (2, 2, 2)) == [(0, 0, 0), (2, 0, 0), (0, 2, 0), (2, 2, 0),
                                                       (0, 0, 2), (2, 0, 2), (0, 2, 2), (2, 2, 2)]


def test_quantize_negative_and_duplicates():
    c = np.array([[0, -0.5, 0.2, 1.9], [0, -1.0, 0.99, 1.0], [0, 3.1, 0, 0], [1, -0.5, 0.2, 1.9],
                  [0, -0.01, 0.5, 1.5]], np.float32)
    for q in (R.quantize_np(c), R.quantize_c(c)):
        assert q.tolist() == [[0, -1, 0, 1], [0, -1, 0, 1], [0, 3, 0, 0], [1, -1, 0, 1], [0, -1, 0, 1]]
    for fn in (R.unique_first_np, R.unique_first_c):
        uc, ui, inv = fn(R.quantize_np(c))
        assert uc.tolist() == [[0, -1, 0, 1], [0, 3, 0, 0], [1, -1, 0, 1]]
        assert ui.tolist() == [0, 2, 3]            # first occurrence, ascending
        assert inv.tolist() == [0, 0, 1, 2, 0]
    f = torch.tensor([[1.0], [3.0], [10.0], [7.0], [5.0]])
    assert R.segment_mean(f, inv, 3).view(-1).tolist() == [3.0, 10.0, 7.0]
    assert R.segment_mean(f, inv, 3, "sum").view(-1).tolist() == [9.0, 10.0, 7.0]


def test_stride_floor_division():
    c = np.array([[0, -1, 0, 3], [0, -2, 1, 4], [0, -3, 2, 5], [1, 7, -8, 0]], np.int32)
    for fn in (R.stride_coords_np, R.stride_coords_c):
        assert fn(c, (2, 2, 2)).tolist() == [[0, -2, 0, 2], [0, -2, 0, 4], [0, -4, 2, 4], [1, 6, -8, 0]]
        assert fn(c, (4, 4, 4)).tolist() == [[0, -4, 0, 0], [0, -4, 0, 4], [0, -4, 0, 4], [1, 4, -8, 0]]


def test_line_conv_shifts():
    # (i) five voxels on the x axis, one-hot weights -> shifts
    coords = np.array([[0, x, 0, 0] for x in range(5)], np.int32)
    offs = R.kernel_offsets((3, 3, 3), (1, 1, 1))
    for fn in (R.kernel_map_np, R.kernel_map_c):
        nbr = fn(coords, coords, offs)
        assert nbr[13].tolist() == [0, 1, 2, 3, 4]
        assert nbr[12].tolist() == [-1, 0, 1, 2, 3]   # offset (-1,0,0): in = out - 1
        assert nbr[14].tolist() == [1, 2, 3, 4, -1]
        assert (np.delete(nbr, [12, 13, 14], axis=0) == -1).all()
    feats = torch.arange(1.0, 6.0).view(5, 1)
    w = torch.zeros(27, 1, 1)
    w[14] = 1.0
    assert R.conv_forward(feats, w, nbr).view(-1).tolist() == [2.0, 3.0, 4.0, 5.0, 0.0]
    pairs = R.pairs_from_dense(nbr)
    assert sorted(pairs) == [12, 13, 14] and pairs[12].tolist() == [[0, 1, 2, 3], [1, 2, 3, 4]]


def test_block_conv_k2s2_and_transpose():
    # (ii)/(iii) a 2x2x2 block through conv k2 s2 and back through convtr k2 s2
    block = np.array([[0, x, y, z] for z in range(2) for y in range(2) for x in range(2)], np.int32)
    mgr = R.OracleManager(block.astype(np.float32))
    ts2 = mgr.stride((1, 1, 1), (2, 2, 2))
    assert mgr.maps[ts2].tolist() == [[0, 0, 0, 0]]
    nbr = mgr.kernel_map((1, 1, 1), ts2, (2, 2, 2))
    assert nbr.reshape(-1).tolist() == list(range(8))          # k = x + 2y + 4z
    feats = torch.arange(1.0, 9.0).view(8, 1)
    w = torch.arange(1.0, 9.0).view(8, 1, 1) * 10
    assert R.conv_forward(feats, w, nbr).item() == sum(10.0 * (i + 1) ** 2 for i in range(8))
    nbr_t = mgr.kernel_map(ts2, (1, 1, 1), (2, 2, 2), transpose=True)
    assert nbr_t.shape == (8, 8)
    up = R.conv_forward(torch.tensor([[2.0]]), w, nbr_t)
    assert up.view(-1).tolist() == [20.0 * (i + 1) for i in range(8)]      # fine voxel f gets in*W[k(f-c)]


def test_kernel1_stride2_keeps_even_lattice():
    # (vi) kernel 1, stride 2: only voxels on the coarse lattice contribute
    coords = np.array([[0, 0, 0, 0], [0, 1, 0, 0], [0, 2, 0, 0], [0, 3, 1, 0]], np.float32)
    mgr = R.OracleManager(coords)
    ts2 = mgr.stride((1, 1, 1), (2, 2, 2))
    assert mgr.maps[ts2].tolist() == [[0, 0, 0, 0], [0, 2, 0, 0]]
    nbr = mgr.kernel_map((1, 1, 1), ts2, (1, 1, 1))
    assert nbr.tolist() == [[0, 2]]
    assert mgr.parents[ts2].tolist() == [0, 0, 1, 1]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_numpy_and_c_restatements_agree(seed):
    rng = np.random.default_rng(seed)
    n = 5000
    c = np.empty((n, 4), np.float32)
    c[:, 0] = rng.integers(0, 3, n)
    c[:, 1:] = rng.uniform(-9, 9, (n, 3))
    q = R.quantize_np(c)
    assert (q == R.quantize_c(c)).all()
    a, b = R.unique_first_np(q), R.unique_first_c(q)
    for x, y in zip(a, b):
        assert (x == y).all()
    uc = a[0]
    assert (uc[a[2]] == q).all() and (np.diff(a[1]) > 0).all()
    for ts_in, k, s in [((1, 1, 1), 3, 1), ((1, 1, 1), 3, 2), ((1, 1, 1), 2, 2), ((1, 1, 1), 1, 2)]:
        out = uc if s == 1 else R.unique_first_np(R.stride_coords_np(uc, (s, s, s)))[0]
        offs = R.kernel_offsets((k,) * 3, ts_in)
        n1, n2 = R.kernel_map_np(uc, out, offs), R.kernel_map_c(uc, out, offs)
        assert (n1 == n2).all()
        # definition check: coord_in == coord_out + offset for every pair
        for kk in (0, len(offs) // 2, len(offs) - 1):
            o = np.nonzero(n1[kk] >= 0)[0]
            assert (uc[n1[kk, o]][:, 1:] == out[o][:, 1:] + np.array(offs[kk])).all()
            assert (uc[n1[kk, o]][:, 0] == out[o][:, 0]).all()
        t = R.transpose_dense(n1, uc.shape[0])
        assert ((t >= 0).sum() == (n1 >= 0).sum())


def test_conv_autograd_gradcheck():
    # (viii) fp64 gradcheck of the oracle convolution on a tiny map
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(0, 4, (40, 4)).astype(np.int32), axis=0)
    c[:, 0] = 0
    c = np.unique(c, axis=0)
    nbr = R.kernel_map_np(c, c, R.kernel_offsets((3, 3, 3), (1, 1, 1)))
    x = torch.randn(c.shape[0], 2, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, 2, 3, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: R.conv_forward(a, b, nbr), (x, w))


def test_plenoxel_decode_oracle_matches_reference_expressions():
    """oracle.plenoxel_decode_np against the reference's own torch expressions for links -> coordinates
    (co3d_3d/src/data/co3d.py:196-203) and the u8 dequantisation (co3d.py:169), evaluated here on the CPU."""
    import torch
    reso = [128, 128, 128]
    rng = np.random.default_rng(0)
    links = torch.from_numpy(np.sort(rng.choice(128 ** 3, size=4000, replace=False)).astype(np.int64))
    coordinates = torch.stack([torch.div(links, (reso[1] * reso[2]), rounding_mode="trunc"),
                               torch.div(links % (reso[1] * reso[2]), reso[2], rounding_mode="trunc"),
                               links % reso[2]], 1).float()
    sh = rng.integers(0, 256, size=(4000, 27), dtype=np.uint8)
    scale, mn = np.float32(2 / 255), np.float32(-1)
    ref_sh = sh.astype(np.float32) * scale + mn
    c, f = R.plenoxel_decode_np(links.numpy(), sh, scale, mn, reso, batch_index=2)
    assert (c[:, 1:] == coordinates.numpy()).all() and (c[:, 0] == 2).all()
    assert (f == ref_sh).all()


def test_iou_counts_oracle_known_answer():
    logits = np.array([[2., 1, 0], [0, 3, 1], [0, 1, 5], [9, 0, 0], [0, 0, 1]], np.float32)
    target = np.array([0, 1, 1, 255, 2])
    out = R.iou_counts_np(logits, target, 3, 255)
    assert out.tolist() == [[1, 2, 1], [1, 1, 1], [1, 1, 2]]


def test_iou_counts_oracle_matches_reference_update_loop():
    """oracle.iou_counts_np against the body of IoUMeter.update (co3d_3d/src/metrics.py:29-41) evaluated with torch."""
    import torch
    g = torch.Generator().manual_seed(3)
    num_classes, ignore_label = 7, 255
    logits = torch.randn(5000, num_classes, generator=g)
    targets = torch.randint(0, num_classes, (5000,), generator=g)
    targets[torch.rand(5000, generator=g) < 0.2] = ignore_label
    preds = logits.argmax(1)
    total_seen, total_correct, total_positive = (torch.zeros(num_classes) for _ in range(3))
    valid = targets != ignore_label
    p, t = preds[valid], targets[valid]
    for i in range(num_classes):
        total_seen[i] += (t == i).sum()
        total_correct[i] += torch.logical_and(t == i, p == t).float().sum()
        total_positive[i] += (p == i).sum()
    out = R.iou_counts_np(logits.numpy(), targets.numpy(), num_classes, ignore_label)
    assert (out[0] == total_seen.numpy()).all() and (out[1] == total_correct.numpy()).all()
    assert (out[2] == total_positive.numpy()).all()


def test_max_pool_oracle_known_answers():
    x = np.array([[1., 5], [3, 2], [0, 9]], np.float32)
    nbr = np.array([[0, -1], [1, 2]])                     # out row 0 <- rows {0, 1}, out row 1 <- row {2}
    out, arg = R.pool_max_np(x, nbr)
    assert out.tolist() == [[3, 5], [0, 9]] and arg.tolist() == [[1, 0], [2, 2]]
    out, arg = R.pool_max_np(x, np.array([[-1], [-1]]))    # no neighbour at all: zeros, arg -1
    assert out.tolist() == [[0, 0]] and arg.tolist() == [[-1, -1]]
    out, arg = R.global_max_np(x, np.array([0, 0, 1]), 3)  # batch 2 is empty
    assert out.tolist() == [[3, 5], [0, 9], [0, 0]] and arg.tolist() == [[1, 0], [2, 2], [-1, -1]]


# ---- next rows f3 / f4: oracle pins ---------------------------------------------------------------------------
def test_instance_norm_oracle_matches_torch_instance_norm():
    """ME composes instance norm from global pooling / broadcast passes; per instance that is exactly
    torch.nn.functional.instance_norm (biased variance) — pin the oracle on it, values and gradients."""
    rng = np.random.default_rng(11)
    batch = np.sort(rng.integers(0, 3, 90))
    x = torch.randn(90, 7, dtype=torch.float64, requires_grad=True)
    w = torch.randn(1, 7, dtype=torch.float64, requires_grad=True)
    b = torch.randn(1, 7, dtype=torch.float64, requires_grad=True)
    out = R.instance_norm(x, batch, 3, w, b, eps=1e-8)
    gy = torch.randn(90, 7, dtype=torch.float64)
    gx, gw, gb = torch.autograd.grad(out, (x, w, b), gy)
    x2 = x.detach().clone().requires_grad_(True)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    parts = []
    for i in range(3):
        rows = torch.from_numpy(np.nonzero(batch == i)[0])
        parts.append(torch.nn.functional.instance_norm(x2[rows].T[None], weight=w2.flatten(), bias=b2.flatten(),
                                                       eps=1e-8)[0].T)
    ref = torch.cat(parts)
    hx, hw, hb = torch.autograd.grad(ref, (x2, w2, b2), gy)
    assert torch.allclose(out, ref, atol=1e-12) and torch.allclose(gx, hx, atol=1e-10)
    assert torch.allclose(gw.flatten(), hw.flatten(), atol=1e-10) and torch.allclose(gb.flatten(), hb.flatten(), atol=1e-10)


def test_seg_head_oracle_known_answer():
    """Two voxels, three points: point 0 and 2 share voxel 1; uniform logits -> loss = log(C); the gradient of a
    shared voxel is the sum over its points; counts follow IoUMeter.update."""
    C = 4
    logits = torch.zeros(2, C, dtype=torch.float64)
    logits[0, 2] = 5.0
    inverse, target = np.array([1, 0, 1]), np.array([0, 2, 3])
    loss, grad, counts = R.seg_head(logits, inverse, target, -100)
    p0 = torch.softmax(logits[0], 0)
    want = (2 * np.log(C) - torch.log(p0[2]).item()) / 3
    assert abs(loss.item() - want) < 1e-12
    g1 = torch.full((C,), 2 * 0.25 / 3, dtype=torch.float64)
    g1[0] -= 1 / 3
    g1[3] -= 1 / 3
    assert torch.allclose(grad[1], g1, atol=1e-12)
    assert counts.tolist() == [[1, 0, 1, 1], [1, 0, 1, 0], [2, 0, 1, 0]]       # argmax of a uniform row = class 0
    # weights: the void (last) class at 0.5 -> weighted mean
    w = torch.tensor([1, 1, 1, 0.5], dtype=torch.float64)
    lw, _, _ = R.seg_head(logits, inverse, target, -100, w)
    assert abs(lw.item() - (np.log(C) - torch.log(p0[2]).item() + 0.5 * np.log(C)) / 2.5) < 1e-12
    # ignored points contribute nothing
    li, gi, ci = R.seg_head(logits, inverse, np.array([0, -100, 3]), -100)
    assert abs(li.item() - np.log(C)) < 1e-12 and float(gi[0].abs().max()) == 0.0 and ci[0].tolist() == [1, 0, 0, 1]


def test_sparse_quantize_oracle_known_answer():
    xyz = np.array([[0.1, 0.1, 0.1], [0.9, 0.2, 0.3], [-0.1, 0.0, 0.0], [1.2, 0.0, 0.0], [0.4, 0.4, 0.4], [1.9, 0.9, 0.9]])
    feats = np.arange(6, dtype=np.float32).reshape(6, 1)
    labels = np.array([3, 3, 1, 2, 5, 2])
    c, f, l, first, inv = R.sparse_quantize_np(xyz, feats, labels, ignore_label=-100, quantization_size=1.0)
    assert c.tolist() == [[0, 0, 0], [-1, 0, 0], [1, 0, 0]]                    # first-occurrence order, floor(-0.1) = -1
    assert first.tolist() == [0, 2, 3] and inv.tolist() == [0, 0, 1, 2, 0, 2]
    assert f.flatten().tolist() == [0.0, 2.0, 3.0]
    assert l.tolist() == [-100, 1, 2]                                          # voxel 0 holds labels 3, 3, 5 -> ignore
    c2, _, _, first2, _ = R.sparse_quantize_np(xyz, quantization_size=0.5)
    assert c2.shape == (5, 3) and first2.tolist() == [0, 1, 2, 3, 5]          # 0.1 and 0.4 share a 0.5-voxel


def test_interpolation_oracle_properties():
    """Trilinear weights: a point's 8 weights sum to one, a lattice point puts all its weight on its own voxel, and a
    linear function stored on a full lattice is reproduced exactly; splat conserves the feature mass."""
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(*[np.arange(-3, 3)] * 3, indexing="ij"), -1).reshape(-1, 3)
    mc = np.concatenate([np.zeros((len(g), 1), np.int64), g], 1).astype(np.int32)
    A, b = np.array([[1.0, 2.0], [0.5, -1.0], [3.0, 0.25]]), np.array([0.3, -2.0])
    f = torch.from_numpy(g @ A + b)
    q = np.concatenate([np.zeros((40, 1)), rng.uniform(-3, 1.99, (40, 3))], 1).astype(np.float32)
    q[0, 1:] = [1.0, -2.0, 0.0]
    corners, rows, w = R.interp_map_np(mc, q)
    assert np.abs(w.sum(0) - 1).max() < 1e-6 and (rows >= 0).all()
    assert w[0, 0] == 1.0 and (w[1:, 0] == 0).all() and mc[rows[0, 0]].tolist() == [0, 1, -2, 0]
    assert corners[5, 7].tolist() == (corners[5, 0] + [0, 1, 1, 1]).tolist()       # corner k = bx + 2 by + 4 bz
    out = R.interpolate(f, rows, w)
    assert np.abs(out.numpy() - (q[:, 1:].astype(np.float64) @ A + b)).max() < 1e-5
    # negative coordinates floor towards -inf; a strided lattice scales the cell
    _, _, w2 = R.interp_map_np(mc, np.array([[0, -0.25, 0, 0]], np.float32))
    assert abs(w2[0, 0] - 0.25) < 1e-7 and abs(w2[1, 0] - 0.75) < 1e-7
    c4, _, w4 = R.interp_map_np(mc, np.array([[0, 3.0, 0, 0]], np.float32), (4, 4, 4))
    assert c4[0, 0].tolist() == [0, 0, 0, 0] and c4[0, 1].tolist() == [0, 4, 0, 0] and abs(w4[1, 0] - 0.75) < 1e-7
    uc, sf, rows_s, _ = R.splat(torch.ones(40, 2, dtype=torch.float64), q)
    assert abs(sf.sum().item() - 80) < 1e-4 and rows_s.max() == uc.shape[0] - 1
