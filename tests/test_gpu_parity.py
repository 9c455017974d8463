"""Parity of the CUDA path (called through the C ABI) against the CPU oracle.

Bars:
  integer outputs (coordinates, unique / inverse maps, kernel maps, pair lists)  : bit-exact
  fp32-faithful conv / BN / pooling vs fp64 oracle                               : |d| <= 1e-4 * (1 + |ref|)
  TF32 tensor-core conv vs fp64 oracle                                            : |d| <= 3e-3 * max|ref|
"""
import numpy as np
import pytest
import torch

from nerf_downstream_b200 import lib as L
from nerf_downstream_b200 import ops, synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
TF32_TOL = 3e-3


def assert_fp32(got: torch.Tensor, ref: torch.Tensor, what=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    bound = FP32_TOL * (1 + ref.abs())
    assert bool((err <= bound).all()), f"{what}: max err {err.max().item():.3e} (ref max {ref.abs().max().item():.3e})"


def assert_tf32(got: torch.Tensor, ref: torch.Tensor, what=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= TF32_TOL * max(scale, 1e-30), f"{what}: max err {err:.3e} vs {TF32_TOL} * {scale:.3e}"


def gpu(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


# ---------------------------------------------------------------------------
# coordinates
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,extent,seed", [(1, 3, 0), (37, 2, 1), (5000, 9, 2), (200_000, 40, 3), (1024, 1, 4)])
def test_quantize_hash_unique_exact(cuda_device, n, extent, seed):
    c, _ = synth.random_cloud(seed, n, extent=extent, n_batch=3)
    cmap, first, inverse, count = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    uc, ui, inv = R.unique_first_np(R.quantize_np(c))
    assert cmap.size == uc.shape[0]
    assert (cmap.coords.cpu().numpy() == uc).all()
    assert (first.cpu().numpy() == ui).all()
    assert (inverse.cpu().numpy() == inv).all()
    assert (count.cpu().numpy() == np.bincount(inv, minlength=uc.shape[0])).all()


def test_quantize_with_tensor_stride(cuda_device):
    c, _ = synth.random_cloud(7, 3000, extent=20, n_batch=2)
    cmap, first, inverse, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (4, 4, 4))
    uc, ui, inv = R.unique_first_np(R.quantize_np(c, (4, 4, 4)))
    assert (cmap.coords.cpu().numpy() == uc).all() and (inverse.cpu().numpy() == inv).all()


def test_empty_and_out_of_range(cuda_device):
    cmap, first, inverse, count = ops.coords_insert(torch.zeros((0, 4), device=cuda_device), L.SRC_FLOAT, (1, 1, 1))
    assert cmap.size == 0 and inverse.numel() == 0
    bad = torch.tensor([[0, 0, 0, 0], [0, 200000.0, 0, 0]], device=cuda_device)
    with pytest.raises(RuntimeError, match="out of the supported range"):
        ops.coords_insert(bad, L.SRC_FLOAT, (1, 1, 1))
    nan = torch.tensor([[0, float("nan"), 0, 0]], device=cuda_device)
    with pytest.raises(RuntimeError, match="out of the supported range"):
        ops.coords_insert(nan, L.SRC_FLOAT, (1, 1, 1))
    edge = torch.tensor([[1022, -131072, 131071, 0], [0, 131071, -131072, 5]], dtype=torch.int32, device=cuda_device)
    cmap, _, _, _ = ops.coords_insert(edge, L.SRC_INT, (1, 1, 1))
    assert cmap.coords.cpu().tolist() == edge.cpu().tolist()


def _maps(cuda_device, seed, n, extent):
    c, f = synth.random_cloud(seed, n, extent=extent, n_batch=2)
    cmap, first, inverse, count = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    uc, ui, inv = R.unique_first_np(R.quantize_np(c))
    return c, f, cmap, uc, inverse, inv


@pytest.mark.parametrize("stride", [2, 4])
def test_stride_map_exact(cuda_device, stride):
    _, _, cmap, uc, _, _ = _maps(cuda_device, 11, 20000, 24)
    ts = (stride,) * 3
    smap, first, parent, count = ops.coords_insert(cmap.coords, L.SRC_STRIDE, ts)
    suc, sui, sinv = R.unique_first_np(R.stride_coords_np(uc, ts))
    assert (smap.coords.cpu().numpy() == suc).all()
    assert (first.cpu().numpy() == sui).all() and (parent.cpu().numpy() == sinv).all()
    # and with the C restatement (ME CPU algorithm structure)
    assert (R.unique_first_c(R.stride_coords_c(uc, ts))[0] == suc).all()


def test_stride_pyramid_matches_level_by_level(cuda_device):
    """coords_insert_pyramid (device-side row counts, one host sync) == one coords_insert per level, and
    its hash tables answer kernel-map probes exactly like the oracle."""
    _, _, cmap, uc, _, _ = _maps(cuda_device, 13, 60000, 40)
    chain = [(2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 16)]
    built = ops.coords_insert_pyramid(cmap, chain)
    src, src_np = cmap, uc
    for ts, (pm, pfirst, pinv, pcount) in zip(chain, built):
        sm, first, inv, count = ops.coords_insert(src.coords, L.SRC_STRIDE, ts)
        assert pm.size == sm.size and (pm.coords == sm.coords).all()
        assert (pfirst == first).all() and (pinv == inv).all() and (pcount == count).all()
        ref_np = R.unique_first_np(R.stride_coords_np(src_np, ts))[0]
        assert (pm.coords.cpu().numpy() == ref_np).all()
        pm.tensor_stride = ts
        offs = ops.kernel_offsets((3, 3, 3), ts, (1, 1, 1))
        km = ops.build_kernel_map(pm, pm, offs)
        assert (km.nbr.cpu().numpy() == R.kernel_map_np(ref_np, ref_np, offs)).all()
        src, src_np = pm, ref_np


KMAP_CASES = [((3, 3, 3), 1), ((3, 3, 3), 2), ((2, 2, 2), 2), ((1, 1, 1), 2), ((5, 5, 5), 1), ((3, 1, 3), 1)]


@pytest.mark.parametrize("ks,stride", KMAP_CASES)
def test_kernel_map_exact(cuda_device, ks, stride):
    _, _, cmap, uc, _, _ = _maps(cuda_device, 21, 30000, 20)
    if stride == 1:
        out_map, out_np = cmap, uc
    else:
        ts = (stride,) * 3
        out_map, _, _, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, ts)
        out_map.tensor_stride = ts
        out_np = R.unique_first_np(R.stride_coords_np(uc, ts))[0]
    offs = ops.kernel_offsets(ks, (1, 1, 1), (1, 1, 1))
    km = ops.build_kernel_map(cmap, out_map, offs)
    ref = R.kernel_map_np(uc, out_np, R.kernel_offsets(ks, (1, 1, 1)))
    assert (km.nbr.cpu().numpy() == ref).all()
    assert (km.tap_count.cpu().numpy() == (ref >= 0).sum(1)).all()
    ref_t = R.transpose_dense(ref, uc.shape[0])
    assert (km.nbr_t.cpu().numpy() == ref_t).all()
    # ME-style pair lists (sparse_conv.py:122-143)
    pairs = km.pairs()
    ref_pairs = R.pairs_from_dense(ref)
    assert sorted(pairs) == sorted(ref_pairs)
    for k in ref_pairs:
        assert pairs[k].dtype == torch.int32 and (pairs[k].cpu().numpy() == ref_pairs[k]).all()
    if km.K <= 32:
        m = km.mask.cpu().numpy().astype(np.uint32)
        for t in range(m.shape[0]):
            want = 0
            for k in range(km.K):
                if (ref[k, t * 128:(t + 1) * 128] >= 0).any():
                    want |= 1 << k
            assert int(m[t]) == want


def test_kernel_map_with_strided_input(cuda_device):
    # 3^3 stride-1 map at tensor stride 2: offsets scale with the input tensor stride
    _, _, cmap, uc, _, _ = _maps(cuda_device, 22, 30000, 24)
    ts = (2, 2, 2)
    m2, _, _, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, ts)
    m2.tensor_stride = ts
    u2 = R.unique_first_np(R.stride_coords_np(uc, ts))[0]
    km = ops.build_kernel_map(m2, m2, ops.kernel_offsets((3, 3, 3), ts, (1, 1, 1)))
    ref = R.kernel_map_c(u2, u2, R.kernel_offsets((3, 3, 3), ts))
    assert (km.nbr.cpu().numpy() == ref).all()
    assert (ref[13] == np.arange(u2.shape[0])).all()


# ---------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------
def _conv_case(cuda_device, seed, n, extent, ks, stride, cin, cout, transpose=False):
    c, _ = synth.random_cloud(seed, n, extent=extent, n_batch=2)
    mgr = R.OracleManager(c)
    cmap, _, _, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    ts_in = (1, 1, 1)
    if stride == 1:
        out_map, ts_out = cmap, ts_in
    else:
        ts_out = (stride,) * 3
        out_map, _, _, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, ts_out)
        out_map.tensor_stride = ts_out
        mgr.stride(ts_in, (stride,) * 3)
    km = ops.build_kernel_map(cmap, out_map, ops.kernel_offsets(ks, ts_in, (1, 1, 1)))
    nbr = mgr.kernel_map(ts_in, ts_out, ks)
    if transpose:
        km = km.swapped()
        nbr = R.transpose_dense(nbr, cmap.size)
    K = len(R.kernel_offsets(ks, ts_in))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(km.m_in, cin, generator=g, dtype=torch.float64)
    w = torch.randn(K, cin, cout, generator=g, dtype=torch.float64) / (K * cin) ** 0.5
    b = torch.randn(1, cout, generator=g, dtype=torch.float64)
    go = torch.randn(km.m_out, cout, generator=g, dtype=torch.float64)
    return km, nbr, x, w, b, go


def _run_conv(cuda_device, km, nbr, x, w, b, go, precision, check):
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ref = R.conv_forward(xr, wr, nbr, br)
    ref.backward(go)
    xg = x.float().to(cuda_device).requires_grad_()
    wg = w.float().to(cuda_device).requires_grad_()
    bg = b.float().to(cuda_device).requires_grad_()
    out = ops.SparseConvFn.apply(xg, wg, bg, km, precision)
    out.backward(go.float().to(cuda_device))
    check(out, ref, "forward")
    check(xg.grad, xr.grad, "dgrad")
    check(wg.grad, wr.grad, "wgrad")
    assert_fp32(bg.grad, br.grad, "dbias")


FP32_CONV_CASES = [
    ((3, 3, 3), 1, 27, 32, False), ((3, 3, 3), 1, 32, 32, False), ((3, 3, 3), 2, 64, 64, False),
    ((2, 2, 2), 2, 32, 32, False), ((2, 2, 2), 2, 48, 40, True), ((1, 1, 1), 2, 64, 128, False),
    ((3, 3, 3), 1, 5, 7, False), ((3, 3, 3), 2, 16, 24, True), ((3, 3, 3), 1, 96, 20, False),
]


@pytest.mark.parametrize("ks,stride,cin,cout,transpose", FP32_CONV_CASES)
def test_conv_fp32_vs_fp64_oracle(cuda_device, ks, stride, cin, cout, transpose):
    case = _conv_case(cuda_device, 31, 6000, 10, ks, stride, cin, cout, transpose)
    _run_conv(cuda_device, *case, L.PREC_FP32, assert_fp32)


TF32_CONV_CASES = [
    ((3, 3, 3), 1, 32, 32, False, 0), ((3, 3, 3), 1, 64, 96, False, 0), ((3, 3, 3), 1, 128, 96, False, 2),
    ((3, 3, 3), 1, 96, 96, False, 1), ((3, 3, 3), 1, 256, 256, False, 0), ((3, 3, 3), 2, 64, 128, False, 0),
    ((2, 2, 2), 2, 32, 32, False, 0), ((2, 2, 2), 2, 256, 128, True, 0), ((1, 1, 1), 2, 128, 256, False, 0),
    ((3, 3, 3), 1, 384, 256, False, 0), ((3, 3, 3), 2, 256, 512, False, 0), ((3, 3, 3), 1, 32, 16, False, 2),
]


@pytest.mark.parametrize("ks,stride,cin,cout,transpose,force_mt", TF32_CONV_CASES)
def test_conv_tf32_tensor_core(cuda_device, ks, stride, cin, cout, transpose, force_mt):
    lib = L.load()
    case = _conv_case(cuda_device, 41, 9000, 11, ks, stride, cin, cout, transpose)
    lib.spc_debug_force_mt(force_mt)
    try:
        _run_conv(cuda_device, *case, L.PREC_TF32, assert_tf32)
    finally:
        lib.spc_debug_force_mt(0)


BF16_TOL = 3e-3   # tensor-core modes within 3e-3 * max|ref| per layer (measured 2.2e-3 .. 2.7e-3)
BF16_CONV_CASES = [((3, 3, 3), 1, 32, 32, False), ((3, 3, 3), 1, 64, 96, False), ((3, 3, 3), 1, 128, 96, False),
                   ((3, 3, 3), 1, 96, 96, False), ((3, 3, 3), 1, 256, 256, False), ((3, 3, 3), 2, 64, 128, False),
                   ((2, 2, 2), 2, 96, 96, False), ((2, 2, 2), 2, 256, 128, True), ((1, 1, 1), 2, 128, 256, False),
                   ((3, 3, 3), 2, 256, 512, False), ((3, 3, 3), 1, 512, 512, False)]   # wgrad in two column groups


@pytest.mark.parametrize("ks,stride,cin,cout,transpose", BF16_CONV_CASES)
def test_conv_bf16_tensor_core(cuda_device, ks, stride, cin, cout, transpose):
    """kind::f16 path on bf16 copies of the rows; stated bound |d| <= 3e-3 * max|ref| vs the fp64 oracle."""
    km, nbr, x, w, b, go = _conv_case(cuda_device, 43, 9000, 11, ks, stride, cin, cout, transpose)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    ref = R.conv_forward(xr, wr, nbr, b)
    ref.backward(go)
    d = cuda_device
    xb, gb = ops.to_bf16(x.float().to(d)), ops.to_bf16(go.float().to(d))
    wg, bg = w.float().to(d), b.float().to(d).view(-1)
    K = wg.shape[0]

    def check(got, want, what):
        err = (got.double().cpu() - want.detach()).abs().max().item()
        scale = want.detach().abs().max().item()
        assert err <= BF16_TOL * scale, f"{what}: max err {err:.3e} vs {BF16_TOL} * {scale:.3e}"
    check(ops.conv_fwd_raw(xb, wg, bg, km, L.PREC_BF16), ref, "forward")
    check(ops.conv_dgrad_raw(gb, wg, km, L.PREC_BF16), xr.grad, "dgrad")
    check(ops.conv_wgrad_raw(xb, gb, km, K, cin, cout, L.PREC_BF16), wr.grad, "wgrad")


def test_conv_tf32_matches_fp32_kernels_large(cuda_device):
    # on-device cross-check at a size the CPU oracle would not finish quickly (~150 K voxels)
    c, _, _ = synth.room_batch(5, 1, 150_000, channels=1)
    cmap, _, _, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(cmap.size, 64, generator=g).to(cuda_device)
    w = (torch.randn(27, 64, 96, generator=g) / (27 * 64) ** 0.5).to(cuda_device)
    go = torch.randn(cmap.size, 96, generator=g).to(cuda_device)
    a = ops.conv_fwd_raw(x, w, None, km, L.PREC_TF32)
    b = ops.conv_fwd_raw(x, w, None, km, L.PREC_FP32)
    assert_tf32(a, b, "fwd")
    assert_tf32(ops.conv_dgrad_raw(go, w, km, L.PREC_TF32), ops.conv_dgrad_raw(go, w, km, L.PREC_FP32), "dgrad")
    # linearity (size-independent property): conv(2x + y) == 2 conv(x) + conv(y)
    y = torch.randn(cmap.size, 64, generator=g).to(cuda_device)
    lhs = ops.conv_fwd_raw(2 * x + y, w, None, km, L.PREC_FP32)
    rhs = 2 * b + ops.conv_fwd_raw(y, w, None, km, L.PREC_FP32)
    assert_fp32(lhs, rhs, "linearity")
    # 3^3 stride-1 maps are symmetric: nbr[k][o] = i  <=>  nbr[26-k][i] = o
    assert bool((km.nbr_t == km.nbr.flip(0)).all())
    # bf16 operands on the same large map (MT = 2 tiles) against the fp32 CUDA-core kernels
    xb, gb = ops.to_bf16(x), ops.to_bf16(go)
    db = ops.conv_dgrad_raw(go, w, km, L.PREC_FP32)

    def close_bf16(got, want, what):
        err = (got - want).abs().max().item()
        assert err <= BF16_TOL * want.abs().max().item(), f"{what}: {err:.3e}"
    close_bf16(ops.conv_fwd_raw(xb, w, None, km, L.PREC_BF16), b, "bf16 fwd")
    close_bf16(ops.conv_dgrad_raw(gb, w, km, L.PREC_BF16), db, "bf16 dgrad")
    dw32 = ops.conv_wgrad_raw(x, go, km, 27, 64, 96, L.PREC_FP32)
    close_bf16(ops.conv_wgrad_raw(xb, gb, km, 27, 64, 96, L.PREC_BF16), dw32, "bf16 wgrad")


# ---------------------------------------------------------------------------
# BN / ReLU / add / pooling / reduction
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("m,C,relu,res", [(1000, 32, False, False), (4097, 27, True, False), (3000, 96, True, True),
                                          (257, 512, False, True), (16, 64, True, False), (50000, 128, True, True)])
def test_batchnorm_matches_torch(cuda_device, m, C, relu, res):
    g = torch.Generator().manual_seed(m + C)
    x = torch.randn(m, C, generator=g, dtype=torch.float64) * 2 + 0.5
    gamma = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(C, generator=g, dtype=torch.float64)
    r = torch.randn(m, C, generator=g, dtype=torch.float64) if res else None
    go = torch.randn(m, C, generator=g, dtype=torch.float64)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    xr, gr, br = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    rr = r.clone().requires_grad_() if res else None
    ref = torch.nn.functional.batch_norm(xr, rm, rv, gr, br, True, 0.1, 1e-5)
    if res:
        ref = ref + rr
    if relu:
        ref = torch.relu(ref)
    ref.backward(go)
    d = cuda_device
    xg, gg, bg = (t.float().to(d).requires_grad_() for t in (x, gamma, beta))
    rg = r.float().to(d).requires_grad_() if res else None
    rmg, rvg = torch.zeros(C, device=d), torch.ones(C, device=d)
    out = ops.BatchNormFn.apply(xg, gg, bg, rmg, rvg, True, 0.1, 1e-5, relu, rg)
    out.backward(go.float().to(d))
    assert_fp32(out, ref, "bn fwd")
    assert_fp32(xg.grad, xr.grad, "bn dx")
    assert_fp32(gg.grad, gr.grad, "bn dgamma")
    assert_fp32(bg.grad, br.grad, "bn dbeta")
    if res:
        assert_fp32(rg.grad, rr.grad, "bn dres")
    assert_fp32(rmg, rm, "running_mean")
    assert_fp32(rvg, rv, "running_var")
    # eval mode uses the running statistics
    ev = ops.BatchNormFn.apply(xg.detach(), gg.detach(), bg.detach(), rmg, rvg, False, 0.1, 1e-5, False, None)
    assert_fp32(ev, torch.nn.functional.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-5), "bn eval")


@pytest.mark.parametrize("m,C,relu,res", [(3000, 96, True, True), (50000, 128, True, False), (777, 32, False, True)])
def test_batchnorm_bf16_side_outputs(cuda_device, m, C, relu, res):
    """bf16 operand mode: BatchNorm apply / backward also write the bf16 copy of y / dx that the next
    convolution consumes, and the backward ReLU mask is read from the bf16 copy.  The fp32 results must equal
    the plain-mode results bit for bit; the side copies must be the RNE rounding of the fp32 rows."""
    g = torch.Generator().manual_seed(m * 3 + C)
    d = cuda_device
    x = (torch.randn(m, C, generator=g) * 2 + 0.5).to(d)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(d), torch.randn(C, generator=g).to(d)
    r = torch.randn(m, C, generator=g).to(d) if res else None
    go = torch.randn(m, C, generator=g).to(d)

    def run():
        xg, gg, bg = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
        rg = r.clone().requires_grad_() if res else None
        out = ops.BatchNormFn.apply(xg, gg, bg, torch.zeros(C, device=d), torch.ones(C, device=d), True, 0.1, 1e-5,
                                    relu, rg)
        grads = torch.autograd.grad(out, [xg, gg, bg] + ([rg] if res else []), go)
        return out, grads

    out_ref, grads_ref = run()
    ops.set_default_precision("bf16")
    try:
        out, grads = run()
        yb = ops._lookup_bf16(out)
        dxb = ops._lookup_bf16(grads[0])
    finally:
        ops.set_default_precision("tf32")
    assert bool((out == out_ref).all())
    for a, b in zip(grads, grads_ref):
        assert bool((a == b).all())
    assert yb is not None and bool((yb == out.to(torch.bfloat16)).all())
    assert dxb is not None and bool((dxb == grads[0].to(torch.bfloat16)).all())


def test_relu_add(cuda_device):
    x = torch.randn(1001, 37, device=cuda_device, requires_grad=True)
    y = ops.ReLUFn.apply(x)
    y.backward(torch.ones_like(y))
    assert torch.equal(y, torch.relu(x.detach())) and torch.equal(x.grad, (x.detach() > 0).float())
    a, b = torch.randn(999, 5, device=cuda_device), torch.randn(999, 5, device=cuda_device)
    assert torch.equal(ops.AddFn.apply(a, b), a + b)


def test_pooling_and_global_pool(cuda_device):
    c, f = synth.random_cloud(51, 20000, extent=16, n_batch=3, channels=64)
    mgr = R.OracleManager(c)
    feats = torch.from_numpy(f)
    x0 = R.segment_mean(feats.double(), mgr.inverse, mgr.maps[(1, 1, 1)].shape[0])
    cmap, first, inverse, count = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    xg = ops.SegmentReduceFn.apply(feats.to(cuda_device), inverse, count, first, cmap.size, 0)
    assert_fp32(xg, x0, "segment mean")
    assert_fp32(ops.SegmentReduceFn.apply(feats.to(cuda_device), inverse, count, first, cmap.size, 1),
                R.segment_mean(feats.double(), mgr.inverse, cmap.size, "sum"), "segment sum")
    ts2 = mgr.stride((1, 1, 1), (2, 2, 2))
    m2, _, parent, cnt2 = ops.coords_insert(cmap.coords, L.SRC_STRIDE, ts2)
    km = ops.build_kernel_map(cmap, m2, ops.kernel_offsets((2, 2, 2), (1, 1, 1), (1, 1, 1)))
    nbr = mgr.kernel_map((1, 1, 1), ts2, (2, 2, 2))
    for avg in (False, True):
        xr = x0.clone().requires_grad_()
        ref = R.sum_pool(xr, nbr, avg)
        xq = xg.detach().clone().requires_grad_()
        out = ops.LocalPoolFn.apply(xq, km, avg)
        go = torch.randn(ref.shape, dtype=torch.float64)
        ref.backward(go)
        out.backward(go.float().to(cuda_device))
        assert_fp32(out, ref, f"pool avg={avg}")
        assert_fp32(xq.grad, xr.grad, f"pool bwd avg={avg}")
    xr = x0.clone().requires_grad_()
    ref = R.global_avg_pool(xr, mgr.maps[(1, 1, 1)], 3)
    xq = xg.detach().clone().requires_grad_()
    out = ops.GlobalPoolFn.apply(xq, cmap.coords, 3, True)
    go = torch.randn(3, 64, dtype=torch.float64)
    ref.backward(go)
    out.backward(go.float().to(cuda_device))
    assert_fp32(out, ref, "global avg")
    assert_fp32(xq.grad, xr.grad, "global avg bwd")
    # slice back to points
    sl = ops.GatherRowsFn.apply(xg, inverse)
    assert_fp32(sl, x0[torch.from_numpy(mgr.inverse.astype(np.int64))], "slice")


# ---------------------------------------------------------------------------
# full-size properties (BASELINE config 2/4 scale: 1 M voxels)
# ---------------------------------------------------------------------------
def test_full_size_round_trips(cuda_device):
    c, _, _ = synth.room_batch(777, 1, 1_000_000, channels=1)
    cg = gpu(c, cuda_device)
    cmap, first, inverse, count = ops.coords_insert(cg, L.SRC_FLOAT, (1, 1, 1))
    assert cmap.size == 1_000_000
    q = torch.floor(cg).int()
    assert bool((cmap.coords[inverse.long()] == q).all())            # inverse map round trip
    assert bool((first[1:] > first[:-1]).all())                      # first-occurrence order
    assert bool((q[first.long()] == cmap.coords).all())
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    ar = torch.arange(cmap.size, device=cuda_device, dtype=torch.int32)
    assert bool((km.nbr[13] == ar).all())                            # centre offset is the identity
    assert bool((km.nbr_t == km.nbr.flip(0)).all())                  # symmetry of 3^3 s1 maps
    # every pair satisfies coord_in == coord_out + offset
    offs = ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1))
    for k in (0, 5, 26):
        o = torch.nonzero(km.nbr[k] >= 0).view(-1)
        d = cmap.coords[km.nbr[k][o].long()] - cmap.coords[o]
        assert bool((d == torch.tensor([0, *offs[k]], device=cuda_device, dtype=torch.int32)).all())
    assert int(km.tap_count.sum()) == int((km.nbr >= 0).sum())
    # stride map: idempotence and parent consistency
    m2, _, parent, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, (2, 2, 2))
    want = cmap.coords.clone()
    want[:, 1:] = torch.div(want[:, 1:], 2, rounding_mode="floor") * 2
    assert bool((m2.coords[parent.long()] == want).all())
    m2b, _, parent_b, _ = ops.coords_insert(m2.coords, L.SRC_STRIDE, (2, 2, 2))
    assert m2b.size == m2.size and bool((parent_b == torch.arange(m2.size, device=cuda_device)).all())


def test_tables_beyond_the_l2_use_pair_buckets_and_stay_exact(cuda_device):
    """More than 2^21 points -> a table of >= 2^22 buckets, where `hash_key` leaves the lowest bit of z out (z pairs
    share a home bucket, csrc/common.cuh).  3 M points with duplicates (two rooms on top of each other): unique /
    inverse / first / count against an on-device restatement built on torch.unique of the packed keys, the self map by
    its properties, and every voxel found again through the general (non-symmetric) probe kernel."""
    c, _, _ = synth.room_batch(5, 1, 2_000_000, channels=1)
    c2, _, _ = synth.room_batch(6, 1, 1_000_000, channels=1)
    pts = torch.cat([gpu(c, cuda_device), gpu(c2, cuda_device)])
    n = pts.shape[0]
    assert L.load().spc_table_slots(n) // 2 >= 1 << 22
    cmap, first, inverse, count = ops.coords_insert(pts, L.SRC_FLOAT, (1, 1, 1))
    q = torch.floor(pts).long()
    key = ((q[:, 0] << 54) | ((q[:, 1] + 131072) << 36) | ((q[:, 2] + 131072) << 18) | (q[:, 3] + 131072))
    uk, inv_sorted, cnt_sorted = torch.unique(key, return_inverse=True, return_counts=True)   # (sorted by key)
    m = uk.numel()
    assert cmap.size == m and m < n                                   # the two rooms share voxels
    # first occurrence of every key, then ranks in first-occurrence order = the row order of the map
    idx = torch.arange(n, device=cuda_device)
    first_sorted = torch.full((m,), n, device=cuda_device, dtype=torch.long).scatter_reduce_(0, inv_sorted, idx, "amin")
    order = torch.argsort(first_sorted)                               # sorted-key id of row r
    rank = torch.empty_like(order)
    rank[order] = torch.arange(m, device=cuda_device)
    assert bool((first.long() == first_sorted[order]).all())
    assert bool((inverse.long() == rank[inv_sorted]).all())
    assert bool((count.long() == cnt_sorted[order]).all())
    assert bool((cmap.coords.long() == q[first.long()]).all())
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    ar = torch.arange(m, device=cuda_device, dtype=torch.int32)
    assert bool((km.nbr[13] == ar).all())
    assert bool((km.nbr_t == km.nbr.flip(0)).all())
    offs = ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1))
    for k in (1, 12, 22):
        o = torch.nonzero(km.nbr[k] >= 0).view(-1)
        d = cmap.coords[km.nbr[k][o].long()] - cmap.coords[o]
        assert bool((d == torch.tensor([0, *offs[k]], device=cuda_device, dtype=torch.int32)).all())
        # ... and no pair is missing: the neighbour exists iff its key is in the set
        want = torch.isin(key[first.long()] + ((offs[k][0] << 36) + (offs[k][1] << 18) + offs[k][2]), uk)
        inside = ((cmap.coords[:, 1:].long() + torch.tensor(offs[k], device=cuda_device)).abs() < 131072).all(1)
        assert bool(((km.nbr[k] >= 0) == (want & inside)).all())
    # the general probe kernel (what stride-2 / transposed maps use) on the same table: every voxel finds itself
    probe = ops.build_kernel_map(cmap, ops.CoordMap(cmap.coords, None, 0, m, (1, 1, 1)), [(0, 0, 0)])
    assert bool((probe.nbr[0] == ar).all())


@pytest.mark.parametrize("n,C", [(1, 20), (777, 20), (50_000, 20), (4096, 51), (100, 3)])
def test_cross_entropy_matches_torch(cuda_device, n, C):
    """fused CE (mean over non-ignored rows) vs torch's fp64 cross_entropy; |d| <= 1e-5 * (1 + |ref|)."""
    g = torch.Generator().manual_seed(n + C)
    logits = (torch.randn(n, C, generator=g) * 3).requires_grad_()
    y = torch.randint(0, C, (n,), generator=g)
    y[torch.rand(n, generator=g) < 0.1] = 255
    if n == 1:
        y[0] = 3
    ref_in = logits.detach().double().requires_grad_()
    ref = torch.nn.functional.cross_entropy(ref_in, y, ignore_index=255)
    ref.backward()
    x = logits.detach().to(cuda_device).requires_grad_()
    loss = ops.cross_entropy(x, y.to(cuda_device), ignore_index=255)
    (loss * 2.5).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * (1 + abs(ref.item()))
    err = (x.grad.double().cpu() - 2.5 * ref_in.grad).abs().max().item()
    assert err <= 1e-6 * (1 + (2.5 * ref_in.grad).abs().max().item()), err


@pytest.mark.parametrize("n,reso,dtype,with_affine", [(1, (128, 128, 128), torch.int64, False),
                                                       (5000, (128, 128, 128), torch.int32, True),
                                                       (200_003, (256, 256, 256), torch.int64, True)])
def test_plenoxel_decode_exact(cuda_device, n, reso, dtype, with_affine):
    """links -> (i,j,k) and u8 SH dequantisation (co3d.py:169,196-203): bit-exact against the numpy restatement."""
    from nerf_downstream_b200 import pipeline
    rng = np.random.default_rng(n)
    links = np.sort(rng.choice(reso[0] * reso[1] * reso[2], size=n, replace=False)).astype(np.int64)
    sh = rng.integers(0, 256, size=(n, 27), dtype=np.uint8)
    scale, mn = np.float32(2.0 / 255.0), np.float32(-1.0)
    aff = None
    if with_affine:
        th = 0.7
        aff = [np.cos(th) * 1.1, 0, np.sin(th) * 1.1, 0, 1.1, 0, -np.sin(th) * 1.1, 0, np.cos(th) * 1.1, 3.25, -7.5, 0.125]
    rc, rf = R.plenoxel_decode_np(links, sh, scale, mn, reso, batch_index=3, affine=aff)
    c, f = pipeline.plenoxel_decode(torch.from_numpy(links).to(dtype).to(cuda_device), torch.from_numpy(sh).to(cuda_device),
                                    float(scale), float(mn), reso, batch_index=3, affine=aff)
    assert (c.cpu().numpy() == rc).all()
    assert (f.cpu().numpy() == rf).all()
    # and the decoded record quantises to the voxels it came from (identity affine)
    if not with_affine:
        cmap, _, _, _ = ops.coords_insert(c, L.SRC_FLOAT, (1, 1, 1))
        assert cmap.size == n


@pytest.mark.parametrize("n,C", [(1, 20), (100_000, 20), (4097, 51)])
def test_seg_metrics_match_reference_loop(cuda_device, n, C):
    """IoUMeter.update (metrics.py:29-41): per-class seen / correct / positive counts, exact; accumulation over calls."""
    from nerf_downstream_b200 import pipeline
    g = torch.Generator().manual_seed(n + C)
    logits = torch.randn(n, C, generator=g)
    y = torch.randint(0, C, (n,), generator=g)
    y[torch.rand(n, generator=g) < 0.15] = 255
    ref = R.iou_counts_np(logits.numpy(), y.numpy(), C, 255)
    meter = pipeline.IoUMeter(C, 255)
    meter.update(logits.to(cuda_device), y.to(cuda_device))
    assert (meter.counts.cpu().numpy() == ref).all()
    meter.update(logits.to(cuda_device), y.to(cuda_device))
    assert (meter.counts.cpu().numpy() == 2 * ref).all()
    miou, ious, macc, accs = meter.compute()
    seen, cor, pos = (ref[i].astype(np.float64) for i in range(3))
    want = np.where(seen > 0, cor / np.maximum(seen + pos - cor, 1), 0.0)
    assert np.allclose(ious.cpu().numpy(), want, atol=1e-6)


def test_row_pitch_inputs(cuda_device):
    """Kernels that take a row pitch read a column slice of a wider tensor in place (what torch.cat's backward hands
    out): bf16 conversion (+ channel padding) and BatchNorm backward must equal the dense-copy results bit for bit."""
    g = torch.Generator().manual_seed(5)
    wide = torch.randn(5000, 128, generator=g).to(cuda_device)
    sl = wide[:, 32:128]  # 96 of 128 columns, pitch 128
    assert not sl.is_contiguous()
    assert bool((ops.to_bf16(sl) == sl.contiguous().to(torch.bfloat16)).all())
    x27 = torch.randn(4097, 27, generator=g).to(cuda_device)
    p = ops.to_bf16(x27, pad_to=32)
    assert p.shape == (4097, 32) and bool((p[:, :27] == x27.to(torch.bfloat16)).all()) and bool((p[:, 27:] == 0).all())
    x = torch.randn(5000, 96, generator=g).to(cuda_device)
    gam, bet = torch.rand(96, generator=g).to(cuda_device) + 0.5, torch.randn(96, generator=g).to(cuda_device)

    def bn_grads(dy):
        xg = x.clone().requires_grad_()
        out = ops.BatchNormFn.apply(xg, gam, bet, torch.zeros(96, device=cuda_device), torch.ones(96, device=cuda_device),
                                    True, 0.1, 1e-5, True, None)
        return torch.autograd.grad(out, xg, dy)[0]
    assert bool((bn_grads(sl) == bn_grads(sl.contiguous())).all())


def test_max_pooling_matches_oracle(cuda_device):
    """MinkowskiMaxPooling (k = 2 and 3, stride 2) and MinkowskiGlobalMaxPooling: values exact, gradient routed to the
    arg-max rows (numpy restatement)."""
    from nerf_downstream_b200 import me as ME
    coords, feats = synth.random_cloud(9, 6000, extent=14, n_batch=3, channels=16)
    x = ME.TensorField(coordinates=gpu(coords, cuda_device), features=gpu(feats, cuda_device)).sparse()
    xf = x.F.detach().cpu().numpy()
    xc = x.C.cpu().numpy()
    for ks in (2, 3):
        xin = ME.SparseTensor(x.F.detach().clone().requires_grad_(), coordinate_map_key=x.coordinate_map_key,
                              coordinate_manager=x.coordinate_manager)
        y = ME.MinkowskiMaxPooling(kernel_size=ks, stride=2, dimension=3)(xin)
        out_c = y.C.cpu().numpy()
        offs = R.kernel_offsets((ks,) * 3, (1, 1, 1))
        nbr = R.kernel_map_np(xc, out_c, offs)
        ref, arg = R.pool_max_np(xf, nbr)
        assert (y.F.detach().cpu().numpy() == ref).all()
        go = torch.randn(y.F.shape, generator=torch.Generator().manual_seed(ks))
        y.F.backward(go.to(cuda_device))
        want = np.zeros_like(xf)
        np.add.at(want, (arg[arg >= 0], np.nonzero(arg >= 0)[1]), go.numpy()[arg >= 0])
        assert np.allclose(xin.F.grad.cpu().numpy(), want, atol=1e-6)
    xin = ME.SparseTensor(x.F.detach().clone().requires_grad_(), coordinate_map_key=x.coordinate_map_key,
                          coordinate_manager=x.coordinate_manager)
    g = ME.MinkowskiGlobalMaxPooling()(xin)
    ref, arg = R.global_max_np(xf, xc[:, 0], 3)
    assert (g.F.detach().cpu().numpy() == ref).all()
    go = torch.randn(3, 16, generator=torch.Generator().manual_seed(1))
    g.F.backward(go.to(cuda_device))
    want = np.zeros_like(xf)
    want[arg, np.arange(16)[None, :].repeat(3, 0)] = go.numpy()
    assert np.allclose(xin.F.grad.cpu().numpy(), want, atol=1e-6)


# ---------------------------------------------------------------------------
# round 2: symmetric self maps, batched weight packing, gradient sinks
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ks,ts", [((3, 3, 3), 1), ((5, 5, 5), 1), ((3, 1, 3), 1), ((3, 3, 3), 2)])
def test_symmetric_self_map_equals_full_probe(cuda_device, ks, ts):
    """spc_kernel_map_sym (half the probes, mirrored writes) == spc_kernel_map, bit for bit, incl. the pair counts."""
    c, _, _ = synth.room_batch(11, 2, 30_000, channels=1)
    cmap, _, _, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    if ts > 1:
        cmap, _, _, _ = ops.coords_insert(cmap.coords, L.SRC_STRIDE, (ts,) * 3)
        cmap.tensor_stride = (ts,) * 3
    offs = ops.kernel_offsets(ks, (ts,) * 3, (1, 1, 1))
    ops.symmetric_maps = False
    try:
        full = ops.build_kernel_map(cmap, cmap, offs)
    finally:
        ops.symmetric_maps = True
    sym = ops.build_kernel_map(cmap, cmap, offs)
    assert bool((full.nbr == sym.nbr).all())
    assert bool((full.tap_count == sym.tap_count).all())
    ref = R.kernel_map_c(cmap.coords.cpu().numpy(), cmap.coords.cpu().numpy(), R.kernel_offsets(ks, (ts,) * 3))
    assert (sym.nbr.cpu().numpy() == ref).all()


def test_batched_weight_packing_and_gradient_sinks(cuda_device):
    """One data-parallel-trainer step == plain autograd + torch SGD: the weight images re-packed by ONE launch after
    the fused SGD kernel (ops.repack_all) and the parameter gradients written straight into the arena by the wgrad /
    BatchNorm kernels (ops.register_grad_sink) change nothing but the launch count."""
    from nerf_downstream_b200 import me as ME
    from nerf_downstream_b200 import trainer
    coords, feats = synth.random_cloud(5, 6000, extent=9, n_batch=2, channels=32)
    c, f = gpu(coords, cuda_device), gpu(feats, cuda_device)

    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(ME.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3), ME.MinkowskiBatchNorm(64),
                                   ME.MinkowskiReLU(), ME.MinkowskiConvolution(64, 32, kernel_size=3, dimension=3)).to(cuda_device)

    ops.set_default_precision("tf32")
    a, b = make(), make()
    tr = trainer.DataParallelTrainer(a, lr=0.05, momentum=0.9, weight_decay=1e-4)
    opt = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    ops.clear_grad_sinks()
    tr2 = None
    for step in range(3):
        # sinks registered for `a` only (clear + re-register keeps `b` on plain autograd accumulation)
        ops.clear_grad_sinks()
        for p in tr.arena.order:
            ops.register_grad_sink(p, lambda param: None)
        before = dict(ops.pack_stats)
        la = a(ME.TensorField(coordinates=c, features=f).sparse()).F.pow(2).mean()
        tr.backward_and_step(la)
        ops.clear_grad_sinks()
        opt.zero_grad()
        lb = b(ME.TensorField(coordinates=c, features=f).sparse()).F.pow(2).mean()
        lb.backward()
        opt.step()
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item()) + 1e-7, (step, la.item(), lb.item())
        if step > 0:
            # `a`'s layers were re-packed by the batched launch after the previous SGD step: every lookup hits
            assert ops.pack_stats["misses"] - before["misses"] <= 4, (before, ops.pack_stats)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=2e-4, atol=2e-6)


PAIR_CASES = [(9001, 32, 32), (20_000, 96, 96), (70_001, 64, 96), (33_000, 128, 128), (12_345, 256, 256),
              (150_000, 128, 96), (200_000, 96, 96)]


@pytest.mark.parametrize("voxels,cin,cout", PAIR_CASES)
def test_cta_pair_conv_kernel_equals_the_single_cta_kernel(cuda_device, voxels, cin, cout):
    """conv_umma_pair.cu (tcgen05 cta_group::2: two CTAs share every weight slab) against conv_umma_kernel on the same
    map: every output row sees the same sequence of MMAs, so forward and dgrad are BIT-identical; the epilogue
    statistics (double atomics in a free order) agree to 1e-12 relative; the accumulate form (dgrad adding into the
    residual branch's gradient) too.  Row counts that are not multiples of the 512-row pair item included."""
    lib = L.load()
    c, _, _ = synth.room_batch(17, 1, voxels, channels=1)
    cmap, _, _, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    km = ops.build_kernel_map(cmap, cmap, ops.kernel_offsets((3, 3, 3), (1, 1, 1), (1, 1, 1)))
    g = torch.Generator(device="cpu").manual_seed(voxels)
    xb = ops.to_bf16(torch.randn(cmap.size, cin, generator=g).to(cuda_device))
    gb = ops.to_bf16(torch.randn(cmap.size, cout, generator=g).to(cuda_device))
    w = (torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5).to(cuda_device)
    base = torch.randn(cmap.size, cin, generator=g).to(cuda_device)

    def run(knob):
        lib.spc_debug_set(8, knob)
        lib.spc_debug_set(9, knob)
        try:
            dw = ops.conv_wgrad_raw(xb, gb, km, 27, cin, cout, L.PREC_BF16)
            out = ops.conv_fwd_raw(xb, w, None, km, L.PREC_BF16)
            sums = torch.zeros(2 * cout, dtype=torch.float64, device=cuda_device)
            out_s, fused = ops.conv_fwd_raw(xb, w, None, km, L.PREC_BF16, bn_sums=sums)
            din = ops.conv_dgrad_raw(gb, w, km, L.PREC_BF16)
            acc = base.clone()
            ops.conv_dgrad_raw(gb, w, km, L.PREC_BF16, add_into=acc)
            torch.cuda.synchronize()
        finally:
            lib.spc_debug_set(8, 0)
            lib.spc_debug_set(9, 0)
        return out, out_s, (sums if fused else None), din, acc, dw

    a, b = run(1), run(2)
    if voxels >= 150_000:     # (on smaller maps the single-CTA kernel splits the offsets over several CTAs: partial sums)
        assert torch.equal(a[0], b[0]), "forward"
        assert torch.equal(a[1], b[1]), "forward with statistics"
        assert torch.equal(a[3], b[3]), "dgrad"
    for i, what in ((0, "forward"), (1, "forward with statistics"), (3, "dgrad")):
        assert (a[i] - b[i]).abs().max() <= 1e-5 * a[i].abs().max(), what
    assert (a[4] - b[4]).abs().max() <= 1e-5 * a[4].abs().max(), "dgrad accumulate"   # (reduce-adds into fp32 rows)
    if a[2] is not None and b[2] is not None:
        assert ((a[2] - b[2]).abs() <= 1e-9 * a[2].abs() + 1e-9).all(), "epilogue statistics"
    if b[2] is not None:     # (the forced pair kernel fuses them on maps where the single-CTA kernel splits offsets)
        o = b[1].double()
        want = torch.cat([o.sum(0), (o * o).sum(0)])
        assert ((b[2] - want).abs() <= 1e-5 * want.abs() + 1e-6 * want.abs().max()).all(), "epilogue statistics vs the rows"
    # wgrad (conv_wgrad_umma_pair.cu): fp32 partial sums meet in dW through red.add in a free order
    assert (a[5] - b[5]).abs().max() <= 1e-4 * a[5].abs().max(), "wgrad"


def test_cta_pair_kernel_on_stride_2_and_transposed_maps(cuda_device):
    """The pair kernel on maps that are not self maps: 2^3 stride-2 convolution (fine -> coarse), its dgrad (coarse ->
    fine: one contribution per fine row) and the transposed convolution on the swapped map (coarse -> fine), K = 8."""
    lib = L.load()
    c, _, _ = synth.room_batch(23, 1, 320_000, channels=1)
    fine, _, _, _ = ops.coords_insert(gpu(c, cuda_device), L.SRC_FLOAT, (1, 1, 1))
    coarse, _, _, _ = ops.coords_insert(fine.coords, L.SRC_STRIDE, (2, 2, 2))
    km = ops.build_kernel_map(fine, coarse, ops.kernel_offsets((2, 2, 2), (1, 1, 1), (1, 1, 1)))
    up = km.swapped()
    g = torch.Generator(device="cpu").manual_seed(3)
    cin, cout = 64, 96
    xb = ops.to_bf16(torch.randn(fine.size, cin, generator=g).to(cuda_device))
    gb = ops.to_bf16(torch.randn(coarse.size, cout, generator=g).to(cuda_device))
    xc = ops.to_bf16(torch.randn(coarse.size, cin, generator=g).to(cuda_device))
    w = (torch.randn(8, cin, cout, generator=g) / (8 * cin) ** 0.5).to(cuda_device)

    def run(knob):
        lib.spc_debug_set(8, knob)
        try:
            down = ops.conv_fwd_raw(xb, w, None, km, L.PREC_BF16)          # [coarse, cout]
            din = ops.conv_dgrad_raw(gb, w, km, L.PREC_BF16)               # [fine, cin]
            upo = ops.conv_fwd_raw(xc, w, None, up, L.PREC_BF16)           # [fine, cout]
            torch.cuda.synchronize()
        finally:
            lib.spc_debug_set(8, 0)
        return down, din, upo

    a, b = run(1), run(2)
    assert a[0].shape == (coarse.size, cout) and a[1].shape == (fine.size, cin) and a[2].shape == (fine.size, cout)
    for i, what in enumerate(("stride-2 forward", "stride-2 dgrad", "transposed forward")):
        assert (a[i] - b[i]).abs().max() <= 1e-5 * a[i].abs().max(), what
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])     # fine-side outputs: large maps, one owner per row
