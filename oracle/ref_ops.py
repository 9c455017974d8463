"""CPU oracle of the sparse-convolution hot path (numpy / torch-CPU / C).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product path (nerf_downstream_b200/) never imports it.

PARITY UNPINNED against MinkowskiEngine: ME (PyPI `MinkowskiEngine`, un-pinned => 0.5.4; see
/root/reference/install.sh:50-52, co3d_3d/README.md:13) is a third-party dependency that is not
vendored under /root/reference and cannot be installed here, and the reference holds no golden
vectors for this path.  What IS pinned:
  * the convolution arithmetic given a kernel map, against the reference's own in-tree PyTorch
    restatement `WeightSparseConvolutionFunction.forward`
    (co3d_3d/src/models/mink/modules/sparse_conv.py:57-152) — fixtures in tests/golden/ generated
    by tests/golden/make_golden.py importing that file in the build container;
  * the kernel-offset numbering, against sparse_conv.py:375-379 (`[4,13,22]` = z axis);
  * BatchNorm against torch.nn.BatchNorm1d (ME wraps it, modules/common.py:24).
Everything else follows SURVEY.md appendix A ([ME-ext] contracts) and hand-checkable known-answer
tests in tests/test_oracle.py.

Two independent restatements of the integer part are kept and cross-checked:
  `*_np`  — vectorised numpy (np.unique / searchsorted on packed keys);
  `*_c`   — oracle_c.c, ME's CPU algorithm structure (sequential hash insert, OpenMP probing).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
C_SRC = HERE / "oracle_c.c"
C_LIB = HERE / "_build" / "liboracle.so"


# ---------------------------------------------------------------------------
# C library
# ---------------------------------------------------------------------------
def build_c(force: bool = False) -> Path:
    """gcc -O2 -fopenmp oracle_c.c -> oracle/_build/liboracle.so"""
    if C_LIB.exists() and not force and C_LIB.stat().st_mtime >= C_SRC.stat().st_mtime:
        return C_LIB
    C_LIB.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", str(C_LIB), str(C_SRC), "-lm"], check=True)
    return C_LIB


_clib = None


def clib():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(str(build_c()))
        P, I64, I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        lib.orc_quantize_f32.argtypes = [P, I64, P, P]
        lib.orc_quantize_f32.restype = None
        lib.orc_stride_coords.argtypes = [P, I64, P, P]
        lib.orc_stride_coords.restype = None
        lib.orc_unique_first.argtypes = [P, I64, P, P, P]
        lib.orc_unique_first.restype = I64
        lib.orc_kernel_map.argtypes = [P, I64, P, I64, P, I, P]
        lib.orc_kernel_map.restype = I
        lib.orc_gather_rows.argtypes = [P, P, I64, I, P]
        lib.orc_gather_rows.restype = None
        _clib = lib
    return _clib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# quantisation / unique (appendix A.2)
# ---------------------------------------------------------------------------
def quantize_np(coords_f: np.ndarray, ts: Sequence[int] = (1, 1, 1)) -> np.ndarray:
    """floor quantisation in float32 then cast: batch floor(b), spatial floor(x/ts)*ts."""
    c = np.asarray(coords_f, dtype=np.float32)
    out = np.empty(c.shape, dtype=np.int32)
    out[:, 0] = np.floor(c[:, 0]).astype(np.int32)
    for a in range(3):
        t = np.float32(ts[a])
        if ts[a] == 1:
            out[:, 1 + a] = np.floor(c[:, 1 + a]).astype(np.int32)
        else:
            out[:, 1 + a] = (np.floor(c[:, 1 + a] / t) * t).astype(np.int32)
    return out


def quantize_c(coords_f: np.ndarray, ts: Sequence[int] = (1, 1, 1)) -> np.ndarray:
    c = np.ascontiguousarray(coords_f, dtype=np.float32)
    out = np.empty(c.shape, dtype=np.int32)
    t = np.asarray(ts, dtype=np.int32)
    clib().orc_quantize_f32(_p(c), c.shape[0], _p(t), _p(out))
    return out


def unique_first_np(coords_i: np.ndarray):
    """(unique rows in first-occurrence order, unique_index, inverse_mapping)."""
    c = np.ascontiguousarray(coords_i, dtype=np.int32)
    if c.shape[0] == 0:
        return c.reshape(0, 4), np.zeros(0, np.int32), np.zeros(0, np.int32)
    _, first, inv = np.unique(c, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    uidx = first[order].astype(np.int32)
    return c[uidx], uidx, rank[inv].astype(np.int32)


def unique_first_c(coords_i: np.ndarray):
    c = np.ascontiguousarray(coords_i, dtype=np.int32)
    n = c.shape[0]
    oc = np.empty((max(n, 1), 4), np.int32)
    ui = np.empty(max(n, 1), np.int32)
    inv = np.empty(max(n, 1), np.int32)
    m = clib().orc_unique_first(_p(c), n, _p(oc), _p(ui), _p(inv))
    if m < 0:
        raise MemoryError("oracle table")
    return oc[:m].copy(), ui[:m].copy(), inv[:n].copy()


def stride_coords_np(coords_i: np.ndarray, ts_out: Sequence[int]) -> np.ndarray:
    c = np.asarray(coords_i, dtype=np.int32)
    out = c.copy()
    for a in range(3):
        out[:, 1 + a] = np.floor_divide(c[:, 1 + a], ts_out[a]) * ts_out[a]
    return out


def stride_coords_c(coords_i: np.ndarray, ts_out: Sequence[int]) -> np.ndarray:
    c = np.ascontiguousarray(coords_i, dtype=np.int32)
    out = np.empty_like(c)
    t = np.asarray(ts_out, dtype=np.int32)
    clib().orc_stride_coords(_p(c), c.shape[0], _p(t), _p(out))
    return out


def segment_mean(feats: torch.Tensor, inverse: np.ndarray, m: int, mode: str = "mean") -> torch.Tensor:
    """UNWEIGHTED_AVERAGE / UNWEIGHTED_SUM reduction of TensorField.sparse() (A.2)."""
    inv = torch.from_numpy(np.asarray(inverse, dtype=np.int64))
    out = torch.zeros((m, feats.shape[1]), dtype=feats.dtype)
    out.index_add_(0, inv, feats)
    if mode == "mean":
        cnt = torch.bincount(inv, minlength=m).clamp(min=1).to(feats.dtype)
        out = out / cnt.unsqueeze(1)
    return out


# ---------------------------------------------------------------------------
# kernel maps (appendix A.4)
# ---------------------------------------------------------------------------
def kernel_offsets(kernel_size: Sequence[int], tensor_stride: Sequence[int], dilation: Sequence[int] = (1, 1, 1)):
    """[(dx,dy,dz)] in ME's index order: first spatial axis fastest; odd sizes centred, even
    sizes start at 0; scaled by the INPUT tensor stride times dilation."""
    ax = []
    for a in range(3):
        k = int(kernel_size[a])
        step = int(tensor_stride[a]) * int(dilation[a])
        lo = -((k - 1) // 2) if k % 2 == 1 else 0
        ax.append([(lo + j) * step for j in range(k)])
    return [(x, y, z) for z in ax[2] for y in ax[1] for x in ax[0]]


_S = 1 << 20
_B = 1 << 19


def _pack(c: np.ndarray) -> np.ndarray:
    c = c.astype(np.int64)
    return ((c[:, 0] * _S + (c[:, 1] + _B)) * _S + (c[:, 2] + _B)) * _S + (c[:, 3] + _B)


def kernel_map_np(in_coords: np.ndarray, out_coords: np.ndarray, offsets) -> np.ndarray:
    """Dense map nbr[K, M_out] (in row or -1) by sorting + binary search."""
    K, m_out = len(offsets), out_coords.shape[0]
    nbr = np.full((K, m_out), -1, dtype=np.int32)
    if in_coords.shape[0] == 0 or m_out == 0:
        return nbr
    keys = _pack(in_coords)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    for k, off in enumerate(offsets):
        q = out_coords.astype(np.int64).copy()
        q[:, 1] += off[0]
        q[:, 2] += off[1]
        q[:, 3] += off[2]
        qk = _pack(q)
        pos = np.searchsorted(skeys, qk)
        pos_c = np.minimum(pos, skeys.shape[0] - 1)
        hit = skeys[pos_c] == qk
        nbr[k, hit] = order[pos_c[hit]].astype(np.int32)
    return nbr


def kernel_map_c(in_coords: np.ndarray, out_coords: np.ndarray, offsets) -> np.ndarray:
    ic = np.ascontiguousarray(in_coords, dtype=np.int32)
    oc = np.ascontiguousarray(out_coords, dtype=np.int32)
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(-1, 3))
    K = off.shape[0]
    nbr = np.empty((K, oc.shape[0]), np.int32)
    rc = clib().orc_kernel_map(_p(ic), ic.shape[0], _p(oc), oc.shape[0], _p(off), K, _p(nbr))
    if rc:
        raise MemoryError("oracle table")
    return nbr


def pairs_from_dense(nbr: np.ndarray) -> Dict[int, np.ndarray]:
    """ME layout {k: int32[2, n_k]} (row 0 in, row 1 out), ascending out row, non-empty k only
    (sparse_conv.py:122-143)."""
    out = {}
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        if o.size:
            out[k] = np.stack([nbr[k, o], o.astype(np.int32)]).astype(np.int32)
    return out


def transpose_dense(nbr: np.ndarray, m_in: int) -> np.ndarray:
    K, m_out = nbr.shape
    t = np.full((K, m_in), -1, np.int32)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        t[k, nbr[k, o]] = o
    return t


# ---------------------------------------------------------------------------
# feature ops (torch CPU; autograd supplies dgrad / wgrad oracles)
# ---------------------------------------------------------------------------
def conv_forward(feats: torch.Tensor, weight: torch.Tensor, nbr: np.ndarray, bias: Optional[torch.Tensor] = None):
    """out[o] += feats[i] @ W[k] for (i,o) in map_k, k ascending; zero-initialised output
    (sparse_conv.py:86,122-143); bias added last (:417-418).  gather -> GEMM -> scatter-add, the
    three-step structure of ME's CPU backend."""
    K, m_out = nbr.shape
    out = torch.zeros((m_out, weight.shape[-1]), dtype=feats.dtype)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if o.size == 0:
            continue
        i = torch.from_numpy(nbr[k, o].astype(np.int64))
        out = out.index_add(0, torch.from_numpy(o.astype(np.int64)), feats.index_select(0, i) @ weight[k])
    if bias is not None:
        out = out + bias.view(1, -1)
    return out


def batch_norm(x, weight, bias, running_mean=None, running_var=None, training=True, momentum=0.1, eps=1e-5):
    """MinkowskiBatchNorm == nn.BatchNorm1d on .F (modules/common.py:24, fcnn.py:138-140)."""
    return torch.nn.functional.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


def sum_pool(feats: torch.Tensor, nbr: np.ndarray, avg: bool = False):
    K, m_out = nbr.shape
    out = torch.zeros((m_out, feats.shape[1]), dtype=feats.dtype)
    cnt = torch.zeros(m_out, dtype=feats.dtype)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if o.size == 0:
            continue
        oi = torch.from_numpy(o.astype(np.int64))
        out = out.index_add(0, oi, feats.index_select(0, torch.from_numpy(nbr[k, o].astype(np.int64))))
        cnt.index_add_(0, oi, torch.ones(o.size, dtype=feats.dtype))
    if avg:
        out = out / cnt.clamp(min=1).unsqueeze(1)
    return out


def global_avg_pool(feats: torch.Tensor, coords: np.ndarray, n_batch: int, avg: bool = True):
    b = torch.from_numpy(coords[:, 0].astype(np.int64))
    out = torch.zeros((n_batch, feats.shape[1]), dtype=feats.dtype).index_add(0, b, feats)
    if avg:
        cnt = torch.bincount(b, minlength=n_batch).clamp(min=1).to(feats.dtype)
        out = out / cnt.unsqueeze(1)
    return out


# ---------------------------------------------------------------------------
# a tiny coordinate manager for the oracle networks
# ---------------------------------------------------------------------------
class OracleManager:
    """Coordinate maps keyed by tensor stride + cached kernel maps (appendix A.3)."""

    def __init__(self, coords_f: np.ndarray, use_c: bool = False):
        self.use_c = use_c
        q = (quantize_c if use_c else quantize_np)(coords_f)
        uc, ui, inv = (unique_first_c if use_c else unique_first_np)(q)
        self.maps: Dict[Tuple[int, int, int], np.ndarray] = {(1, 1, 1): uc}
        self.unique_index = ui
        self.inverse = inv
        self.parents: Dict[Tuple[int, int, int], np.ndarray] = {}
        self.kmaps: Dict[tuple, np.ndarray] = {}

    def stride(self, ts_in, stride):
        ts_out = tuple(a * b for a, b in zip(ts_in, stride))
        if ts_out == tuple(ts_in):
            return ts_out
        if ts_out not in self.maps:
            sc = (stride_coords_c if self.use_c else stride_coords_np)(self.maps[tuple(ts_in)], ts_out)
            uc, _, inv = (unique_first_c if self.use_c else unique_first_np)(sc)
            self.maps[ts_out] = uc
            self.parents[ts_out] = inv
        return ts_out

    def kernel_map(self, ts_in, ts_out, kernel_size, dilation=(1, 1, 1), transpose=False):
        key = (tuple(ts_in), tuple(ts_out), tuple(kernel_size), tuple(dilation), transpose)
        if key not in self.kmaps:
            if transpose:
                fwd = self.kernel_map(ts_out, ts_in, kernel_size, dilation, False)  # fine -> coarse
                self.kmaps[key] = transpose_dense(fwd, self.maps[tuple(ts_out)].shape[0])
            else:
                offs = kernel_offsets(kernel_size, ts_in, dilation)
                fn = kernel_map_c if self.use_c else kernel_map_np
                self.kmaps[key] = fn(self.maps[tuple(ts_in)], self.maps[tuple(ts_out)], offs)
        return self.kmaps[key]


# ---------------------------------------------------------------------------
# the two ends of the path
# ---------------------------------------------------------------------------
def plenoxel_decode_np(links: np.ndarray, sh_u8: np.ndarray, sh_scale: float, sh_min: float, reso, batch_index=0,
                       affine=None):
    """co3d_3d/src/data/co3d.py:196-203 (links -> (i,j,k), trunc division) and :169 (sh.astype(float32) * scale +
    min); `affine` (12 floats) = 3x3 then translation, products and sums rounded in float32 left to right."""
    links = links.astype(np.int64)
    r1, r2 = int(reso[1]), int(reso[2])
    x = (links // (r1 * r2)).astype(np.float32)
    y = ((links % (r1 * r2)) // r2).astype(np.float32)
    z = (links % r2).astype(np.float32)
    if affine is not None:
        a = np.asarray(affine, np.float32)

        def row(i):
            acc = (a[3 * i] * x).astype(np.float32)
            acc = (acc + (a[3 * i + 1] * y).astype(np.float32)).astype(np.float32)
            acc = (acc + (a[3 * i + 2] * z).astype(np.float32)).astype(np.float32)
            return (acc + a[9 + i]).astype(np.float32)
        x, y, z = row(0), row(1), row(2)
    coords = np.stack([np.full_like(x, np.float32(batch_index)), x, y, z], 1).astype(np.float32)
    feats = (sh_u8.astype(np.float32) * np.float32(sh_scale)).astype(np.float32) + np.float32(sh_min)
    return coords, feats.astype(np.float32)


def random_crop_select_np(xyz: np.ndarray, u3, size3):
    """RandomCrop.__call__ for ONE draw `u3` of the box position (co3d_3d/src/data/transforms.py:206-226): returns
    (indices of the kept rows in their order, fits) — `fits`: the box covers the extent on every axis, the reference then
    returns its input unchanged (:217-218).  float32 arithmetic in the reference's order of operations."""
    c = np.asarray(xyz, np.float32)
    mn = c.min(0, keepdims=True)
    norm = (c - mn).astype(np.float32)
    max_coords = norm.max(0, keepdims=True)
    size = np.asarray(size3, np.float32).reshape(1, 3)
    rng = np.clip((max_coords - size).astype(np.float32), 0, np.inf).astype(np.float32)
    fits = bool(np.prod(rng == 0))
    lo = (np.asarray(u3, np.float32).reshape(1, 3) * rng).astype(np.float32)
    hi = (lo + size).astype(np.float32)
    sel = np.logical_and(np.prod(norm > lo, 1), np.prod(norm < hi, 1)).astype(bool)
    return np.nonzero(sel)[0].astype(np.int32), fits


def iou_counts_np(logits: np.ndarray, target: np.ndarray, num_classes: int, ignore_label: int) -> np.ndarray:
    """IoUMeter.update (co3d_3d/src/metrics.py:29-41) on preds = argmax(logits): [3, C] seen / correct / positive."""
    preds = logits.argmax(1)
    valid = target != ignore_label
    preds, target = preds[valid], target[valid]
    out = np.zeros((3, num_classes), np.int64)
    for i in range(num_classes):
        out[0, i] = (target == i).sum()
        out[1, i] = np.logical_and(target == i, preds == target).sum()
        out[2, i] = (preds == i).sum()
    return out


def pool_max_np(x: np.ndarray, nbr: np.ndarray):
    """Max pooling over a kernel-map region (ME.MinkowskiMaxPooling semantics: maximum over the EXISTING neighbours
    of each output voxel): returns (out [M_out, C], arg [M_out, C] = winning input row, -1 when there is none)."""
    K, m_out = nbr.shape
    C = x.shape[1]
    out = np.zeros((m_out, C), x.dtype)
    arg = np.full((m_out, C), -1, np.int64)
    for o in range(m_out):
        rows = nbr[:, o][nbr[:, o] >= 0]
        if rows.size:
            vals = x[rows]                      # [n, C], in offset order
            w = vals.argmax(0)                  # first maximum = lowest offset index
            out[o] = vals[w, np.arange(C)]
            arg[o] = rows[w]
    return out, arg


def global_max_np(x: np.ndarray, batch: np.ndarray, n_batch: int):
    """Per-batch-index maximum (ME.MinkowskiGlobalMaxPooling): (out [B, C], arg [B, C] = first row attaining it)."""
    C = x.shape[1]
    out = np.zeros((n_batch, C), x.dtype)
    arg = np.full((n_batch, C), -1, np.int64)
    for b in range(n_batch):
        rows = np.nonzero(batch == b)[0]
        if rows.size:
            w = x[rows].argmax(0)
            out[b] = x[rows][w, np.arange(C)]
            arg[b] = rows[w]
    return out, arg


# ---------------------------------------------------------------------------
# next rows f3 / f4: segmentation head, instance norm, sparse_quantize
# ---------------------------------------------------------------------------
def seg_head(voxel_logits: torch.Tensor, inverse: Optional[np.ndarray], target: np.ndarray, ignore_index: int,
             weight: Optional[torch.Tensor] = None):
    """The reference's three separate steps on the CPU, in its own torch expressions:
      logits = out.slice(x).F                       — voxel rows gathered through the inverse map (res16unet.py:435)
      loss   = F.cross_entropy(logits, labels, weight=, ignore_index=)       (segmentation_training.py:35-44)
      counts = IoUMeter.update(logits.argmax(1), labels)                     (metrics.py:29-41)
    Returns (loss, d loss / d voxel_logits, counts [3, C]) in the dtype of `voxel_logits` (use float64)."""
    v = voxel_logits.detach().clone().requires_grad_(True)
    idx = torch.arange(v.shape[0]) if inverse is None else torch.from_numpy(np.asarray(inverse)).long()
    logits = v[idx]
    t = torch.from_numpy(np.asarray(target)).long()
    loss = torch.nn.functional.cross_entropy(logits, t, weight=weight, ignore_index=ignore_index)
    (grad,) = torch.autograd.grad(loss, v)
    counts = iou_counts_np(logits.detach().numpy(), t.numpy(), v.shape[1], ignore_index)
    return loss.detach(), grad, counts


def instance_norm(x: torch.Tensor, batch: np.ndarray, n_batch: int, weight: Optional[torch.Tensor] = None,
                  bias: Optional[torch.Tensor] = None, eps: float = 1e-8) -> torch.Tensor:
    """ME.MinkowskiInstanceNorm as ME composes it [ME-ext, MinkowskiNormalization.py]: per batch index,
    mean = avg(x); centered = x - mean; var = avg(centered^2); out = centered / sqrt(var + eps); then the module's
    `* weight + bias` ([1, C]).  Differentiable torch expressions: autograd of this function is the gradient oracle."""
    b = torch.from_numpy(np.asarray(batch)).long()
    out = torch.zeros_like(x)
    for i in range(n_batch):
        rows = torch.nonzero(b == i).flatten()
        if rows.numel() == 0:
            continue
        xi = x[rows]
        centered = xi - xi.mean(0, keepdim=True)
        var = (centered ** 2).mean(0, keepdim=True)
        out = out.index_copy(0, rows, centered / torch.sqrt(var + eps))
    if weight is not None:
        out = out * weight.reshape(1, -1)
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return out


def sparse_quantize_np(coordinates: np.ndarray, features=None, labels=None, ignore_label: int = -100,
                       quantization_size=None):
    """ME.utils.sparse_quantize [ME-ext] as the reference calls it (scannet.py:235-242), voxels in first-occurrence
    order: (discrete coords [M, 3] int32, features[first], voxel labels (ignore_label where the points of a voxel
    disagree), first index [M], inverse [N])."""
    c = np.asarray(coordinates)
    if quantization_size is not None:
        c = np.floor(c / quantization_size)
    elif np.issubdtype(c.dtype, np.floating):
        c = np.floor(c)
    disc = c.astype(np.int32)
    _, first, inverse = np.unique(disc, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    order = np.argsort(first, kind="stable")            # np.unique sorts lexicographically; we want first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    first, inverse = first[order], rank[inverse]
    out_labels = None
    if labels is not None:
        labels = np.asarray(labels).astype(np.int64)
        out_labels = labels[first].copy()
        for j in range(labels.shape[0]):                 # sequential, like ME's quantize_label
            if labels[j] != labels[first[inverse[j]]]:
                out_labels[inverse[j]] = ignore_label
        out_labels = out_labels.astype(np.int32)
    return disc[first], (None if features is None else np.asarray(features)[first]), out_labels, first, inverse


def interp_map_np(map_coords: np.ndarray, query: np.ndarray, ts: Sequence[int] = (1, 1, 1)):
    """Trilinear interpolation map of ME.MinkowskiInterpolation / TensorField.splat [ME-ext]: for every query point
    (b, x, y, z) the 8 lattice voxels around it, corner k = bx + 2 by + 4 bz at floor(x / ts) * ts + bit * ts with
    weight prod_axis (bit ? f : 1 - f), f = x / ts - floor(x / ts) (float32 arithmetic, like the kernel).
    Returns (corner coordinates int32 [N, 8, 4], rows in `map_coords` int64 [8, N] (-1 = absent), weights f32 [8, N])."""
    q = np.asarray(query, np.float32)
    n = q.shape[0]
    tsf = np.asarray(ts, np.float32)
    t = (q[:, 1:] / tsf).astype(np.float32)
    fl = np.floor(t)
    frac = (t - fl).astype(np.float32)
    lower = np.concatenate([np.floor(q[:, :1]).astype(np.int32), fl.astype(np.int32) * np.asarray(ts, np.int32)], 1)
    corners = np.repeat(lower[:, None, :], 8, 1)
    w = np.ones((8, n), np.float32)
    for k in range(8):
        for a in range(3):
            bit = (k >> a) & 1
            corners[:, k, 1 + a] += bit * int(ts[a])
            w[k] = (w[k] * (frac[:, a] if bit else (np.float32(1) - frac[:, a]))).astype(np.float32)
    lut = {tuple(c): i for i, c in enumerate(np.asarray(map_coords).tolist())}
    rows = np.full((8, n), -1, np.int64)
    for j in range(n):
        for k in range(8):
            rows[k, j] = lut.get(tuple(corners[j, k].tolist()), -1)
    return corners, rows, w


def interpolate(feats: torch.Tensor, rows: np.ndarray, w: np.ndarray) -> torch.Tensor:
    """out[j] = sum_k w[k, j] * feats[rows[k, j]] over the present corners (differentiable torch expression)."""
    out = torch.zeros((rows.shape[1], feats.shape[1]), dtype=feats.dtype)
    for k in range(rows.shape[0]):
        ok = torch.from_numpy(rows[k] >= 0)
        r = torch.from_numpy(np.where(rows[k] >= 0, rows[k], 0))
        out = out + torch.where(ok[:, None], feats[r] * torch.from_numpy(w[k]).to(feats.dtype)[:, None],
                                torch.zeros((), dtype=feats.dtype))
    return out


def splat(feats: torch.Tensor, query: np.ndarray, ts: Sequence[int] = (1, 1, 1)):
    """TensorField.splat: (voxel coordinates int32 [M, 4] in order of first touch (point-major, corner-minor),
    voxel features [M, C] = sum over (point, corner) of w * feats[point])."""
    corners, _, w = interp_map_np(np.zeros((0, 4), np.int32), query, ts)
    flat = corners.reshape(-1, 4)
    uc, _, inv = unique_first_np(flat)
    rows = np.asarray(inv).reshape(-1, 8).T                      # [8, N]
    out = torch.zeros((uc.shape[0], feats.shape[1]), dtype=feats.dtype)
    for k in range(8):
        out = out.index_add(0, torch.from_numpy(rows[k]).long(), feats * torch.from_numpy(w[k]).to(feats.dtype)[:, None])
    return uc, out, rows, w
