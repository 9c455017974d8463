"""Functional layer over the C ABI: coordinate maps, kernel maps and autograd Functions.

Host-side mirror of what MinkowskiEngine's `CoordinateMapManager` + conv/pool/BN functions do
for the reference's call sites.  All arithmetic happens in
`libsparseconv_b200.so`; torch is used for memory, streams and autograd bookkeeping only.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Optional, Sequence, Tuple

import torch

from . import lib as L

_I3 = ctypes.c_int32 * 3

# precision used by convolution layers unless a layer overrides it
PRECISIONS = {"tf32": L.PREC_TF32, "fp32": L.PREC_FP32, "bf16": L.PREC_BF16}
_default_precision = L.PREC_TF32


def set_default_precision(mode: str) -> None:
    """'tf32' (tcgen05 kind::tf32 on the fp32 rows, default), 'bf16' (tcgen05 kind::f16 on bf16 copies of
    the rows: half the gather bytes, twice the MMA rate, fp32 accumulation) or 'fp32' (CUDA-core FFMA)."""
    global _default_precision
    _default_precision = PRECISIONS[mode]


def default_precision() -> int:
    return _default_precision


def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


# Scratch for the kernels (packed weight slabs, BN partial sums): one grow-only buffer per (device,
# stream).  Kernels on one stream run in order, so the next call may overwrite it; nothing in it is
# read after the call that filled it.
_workspaces: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, L.stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ---------------------------------------------------------------------------
# optional per-kernel-class timing (bench.py roofline); off unless a profiler is installed
# ---------------------------------------------------------------------------
class KernelProfiler:
    """CUDA-event timing of every library call on the launching stream, grouped by kernel class,
    with the ALGORITHMIC flops / bytes of each call (formulas: SURVEY.md §8d)."""

    def __init__(self):
        self.records = []

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        return e

    def end(self, name, e0, flops=0.0, nbytes=0.0, detail=""):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream())
        self.records.append((name, e0, e1, float(flops), float(nbytes), detail))

    def summary_detail(self):
        """{(class, detail): {...}} — per layer shape."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, fl, by, detail in self.records:
            d = out.setdefault((name, detail), {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
            d["flops"] += fl
            d["bytes"] += by
        return out

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, fl, by, _ in self.records:
            d = out.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
            d["flops"] += fl
            d["bytes"] += by
        return out


_profiler: Optional[KernelProfiler] = None


def set_profiler(p: Optional[KernelProfiler]) -> None:
    global _profiler
    _profiler = p


# ---------------------------------------------------------------------------
# coordinate maps
# ---------------------------------------------------------------------------
class CoordMap:
    """One coordinate map: unique int32 rows (b,x,y,z) + the hash table that indexes them."""

    __slots__ = ("coords", "table", "n_slots", "size", "tensor_stride", "n_batch_cache")

    def __init__(self, coords, table, n_slots, size, tensor_stride):
        self.coords = coords            # int32 [M,4]
        self.table = table              # uint8 [n_slots*16]
        self.n_slots = n_slots
        self.size = size
        self.tensor_stride = tuple(int(t) for t in tensor_stride)
        self.n_batch_cache = None


def coords_insert(src: torch.Tensor, kind: int, ts: Sequence[int]):
    """Quantise + hash + unique.  Returns (CoordMap, first_idx, inverse, count).

    first_idx[r] is ME's `unique_index`, inverse[j] ME's `inverse_mapping` (both int32 here).
    One host synchronisation: the number of unique rows is read back to size the outputs.
    """
    lib = L.load()
    if src.dim() != 2 or src.shape[1] != 4:
        raise RuntimeError(f"coordinates must be [N,4] (batch,x,y,z), got {tuple(src.shape)}")
    want = torch.float32 if kind == L.SRC_FLOAT else torch.int32
    if src.dtype != want:
        raise RuntimeError(f"coordinate dtype {src.dtype} does not match source kind {kind}")
    src = src.contiguous()
    dev = src.device
    n = src.shape[0]
    n_slots = int(lib.spc_table_slots(n))
    table = _empty(n_slots * L.SLOT_BYTES, torch.uint8, dev)
    coords = _empty((n, 4), torch.int32, dev)
    first = _empty(n, torch.int32, dev)
    inverse = _empty(n, torch.int32, dev)
    count = _empty(n, torch.int32, dev)
    status = _empty(2, torch.int32, dev)
    ws_bytes = int(lib.spc_coords_insert_workspace(n))
    ws = _empty(ws_bytes, torch.uint8, dev)
    ts_arr = _I3(*[int(t) for t in ts])
    e0 = _profiler.begin() if _profiler else None
    L.check(lib.spc_coords_insert(L.ptr(src), n, kind, ctypes.cast(ts_arr, ctypes.c_void_p), L.ptr(table),
                                  n_slots, L.ptr(coords), L.ptr(first), L.ptr(inverse), L.ptr(count),
                                  L.ptr(status), L.ptr(ws), ws_bytes, L.stream()), "spc_coords_insert")
    if e0 is not None:
        _profiler.end("coords_insert(hash)", e0, 0, 20.0 * n + 36.0 * n, f"N{n}")
    m, err = status.tolist()  # host sync
    if err:
        raise RuntimeError("coordinate out of the supported range: batch index must be in [0,1022] and "
                           "x,y,z in [-131072,131071] (and finite)")
    cmap = CoordMap(coords[:m], table, n_slots, m, ts)
    return cmap, first[:m], inverse, count[:m]


def coords_insert_pyramid(in_map: CoordMap, ts_list):
    """Stride maps ts_list[0] <- in_map, ts_list[1] <- ts_list[0], ... enqueued back to back with
    upper-bound allocations (a strided map never has more rows than its parent) and device-side row
    counts, then ONE host synchronisation for all levels instead of one per level.
    Returns [(CoordMap, first_idx, inverse, count), ...] exactly as coords_insert would."""
    lib = L.load()
    dev = in_map.coords.device
    n_up = in_map.size
    src, n_dev = in_map.coords, None
    pend = []
    e0 = _profiler.begin() if _profiler else None
    for ts in ts_list:
        n_slots = int(lib.spc_table_slots(n_up))
        table = _empty(n_slots * L.SLOT_BYTES, torch.uint8, dev)
        coords = _empty((n_up, 4), torch.int32, dev)
        first = _empty(n_up, torch.int32, dev)
        inverse = _empty(n_up, torch.int32, dev)
        count = _empty(n_up, torch.int32, dev)
        status = _empty(2, torch.int32, dev)
        ws_bytes = int(lib.spc_coords_insert_workspace(n_up))
        ws = _empty(ws_bytes, torch.uint8, dev)
        ts_arr = _I3(*[int(t) for t in ts])
        L.check(lib.spc_coords_insert_dev(L.ptr(src), n_up, L.ptr(n_dev), L.SRC_STRIDE,
                                          ctypes.cast(ts_arr, ctypes.c_void_p), L.ptr(table), n_slots, L.ptr(coords),
                                          L.ptr(first), L.ptr(inverse), L.ptr(count), L.ptr(status), L.ptr(ws),
                                          ws_bytes, L.stream()), "spc_coords_insert_dev")
        pend.append((ts, table, n_slots, coords, first, inverse, count, status))
        src, n_dev = coords, status
    if e0 is not None:
        _profiler.end("coords_insert(hash)", e0, 0, (20.0 + 36.0) * n_up * len(ts_list), f"pyramid x{len(ts_list)} N{n_up}")
    sizes = torch.stack([p[7] for p in pend]).tolist()  # the one host sync
    out, n_prev = [], in_map.size
    for (ts, table, n_slots, coords, first, inverse, count, _), (m, err) in zip(pend, sizes):
        if err:
            raise RuntimeError("coordinate out of the supported range: batch index must be in [0,1022] and "
                               "x,y,z in [-131072,131071]")
        out.append((CoordMap(coords[:m], table, n_slots, m, ts), first[:m], inverse[:n_prev], count[:m]))
        n_prev = m
    return out


def kernel_offsets(kernel_size: Sequence[int], tensor_stride: Sequence[int], dilation: Sequence[int]):
    """HYPER_CUBE offsets; index order: first spatial axis fastest (sparse_conv.py:375-379).

    Odd kernel sizes are centred, even ones start at 0; offsets are scaled by the INPUT tensor
    stride times the dilation."""
    ks = [int(k) for k in kernel_size]
    axes = []
    for a in range(3):
        k, step = ks[a], int(tensor_stride[a]) * int(dilation[a])
        lo = -(k - 1) // 2 if k % 2 == 1 else 0
        axes.append([(lo + j) * step for j in range(k)])
    offs = []
    for jz in range(ks[2]):
        for jy in range(ks[1]):
            for jx in range(ks[0]):
                offs.append((axes[0][jx], axes[1][jy], axes[2][jz]))
    return offs


class KernelMap:
    """Dense offset-major kernel map nbr[K, M_out] (+ lazily its transpose, masks, pair lists)."""

    def __init__(self, nbr, tap_count, K, m_in, m_out):
        self.nbr = nbr
        self.tap_count = tap_count
        self.K = K
        self.m_in = m_in
        self.m_out = m_out
        self._mask = None
        self._nbr_t = None
        self._mask_t = None
        self._pairs = None
        self._n_pairs = None
        # centrally symmetric self map (odd kernel, stride 1): nbr_t[k] == nbr[K-1-k], so dgrad runs on `nbr` with the
        # kernel offsets reversed in the packed weights and the transposed map is never built
        self.symmetric = False

    @property
    def n_pairs(self) -> int:
        """Number of (in,out) pairs; host sync, used for reporting only."""
        if self._n_pairs is None:
            self._n_pairs = int(self.tap_count.sum().item()) if self.tap_count is not None else int(self.m_out)
        return self._n_pairs

    @property
    def mask(self):
        if self._mask is None and self.K <= 32:
            self._mask = tile_mask(self.nbr, self.m_out, self.K)
        return self._mask

    @property
    def nbr_t(self):
        if self._nbr_t is None:
            lib = L.load()
            t = _empty((self.K, self.m_in), torch.int32, self.nbr.device)
            L.check(lib.spc_kernel_map_transpose(L.ptr(self.nbr), self.m_out, self.m_in, self.K, L.ptr(t),
                                                 L.stream()), "spc_kernel_map_transpose")
            self._nbr_t = t
        return self._nbr_t

    @property
    def mask_t(self):
        if self._mask_t is None and self.K <= 32:
            self._mask_t = tile_mask(self.nbr_t, self.m_in, self.K)
        return self._mask_t

    def masked(self, bits: Optional[int], transposed: bool = False):
        """Tile mask restricted to the kernel offsets in `bits` (bit k = offset k is kept; None = all): pruned
        offsets of a weight-sparse convolution are skipped by the tensor-core kernels like offsets without a
        neighbour in the tile."""
        m = self.mask_t if transposed else self.mask
        if bits is None or m is None:
            return m
        key = (int(bits), bool(transposed))
        cache = self.__dict__.setdefault("_masked", {})
        v = cache.get(key)
        if v is None:
            b = int(bits) & 0xFFFFFFFF
            v = cache[key] = torch.bitwise_and(m, b - (1 << 32) if b >= (1 << 31) else b)
        return v

    def swapped(self) -> "KernelMap":
        """The same pairs with in/out roles exchanged (transposed convolution)."""
        km = KernelMap(self.nbr_t, self.tap_count, self.K, self.m_out, self.m_in)
        km._n_pairs = self._n_pairs
        km._mask = self.mask_t
        km._nbr_t = self.nbr
        km._mask_t = self.mask
        return km

    def pairs(self):
        """ME-style dict {k: IntTensor[2, n_k]} (row 0 = in rows, row 1 = out rows, ascending out
        row), only non-empty offsets (sparse_conv.py:122-143)."""
        if self._pairs is None:
            lib = L.load()
            dev = self.nbr.device
            counts = self.tap_count.tolist()
            total = int(sum(counts))
            pairs = _empty((2, max(total, 1)), torch.int32, dev)
            tap_off = _empty(self.K + 1, torch.int32, dev)
            ws_bytes = int(lib.spc_pairs_workspace(self.m_out, self.K))
            ws = _empty(ws_bytes, torch.uint8, dev)
            L.check(lib.spc_kernel_map_pairs(L.ptr(self.nbr), self.m_out, self.K, max(total, 1), L.ptr(pairs),
                                             L.ptr(tap_off), L.ptr(ws), ws_bytes, L.stream()),
                    "spc_kernel_map_pairs")
            out, start = {}, 0
            for k, c in enumerate(counts):
                if c > 0:
                    out[k] = pairs[:, start:start + c]
                start += c
            self._pairs = out
        return self._pairs


def tile_mask(nbr, m, K):
    lib = L.load()
    n_tiles = (m + 127) // 128
    mask = _empty(max(n_tiles, 1), torch.int32, nbr.device)
    L.check(lib.spc_tile_mask(L.ptr(nbr), m, K, L.ptr(mask), L.stream()), "spc_tile_mask")
    return mask


hollow_rows = True           # bf16 mode: BatchNorm outputs that only feed convolutions keep their fp32 rows unwritten
recompute_relu_mask = True   # BatchNorm backward re-computes the mask of a residual-free ReLU from x
fuse_conv_bn = True          # ME surface: convolution + BatchNorm (+ residual, ReLU) as one autograd node in bf16 mode
lazy_cat = True              # bf16 mode: ME.cat assembles the bf16 operand copies, the fp32 concatenation stays hollow
symmetric_dgrad = True       # dgrad of a symmetric self map reads the forward map with reversed offsets (no transposed map)
symmetric_maps = True   # build self maps of odd stride-1 kernels with spc_kernel_map_sym (tests switch it off to compare)


_offset_cache: dict = {}   # offsets tuple -> (ctypes int32 [3K] array, centrally symmetric?)  (a few dozen entries per network)


def _offsets_ctypes(offsets):
    key = tuple(tuple(int(v) for v in off) for off in offsets)
    hit = _offset_cache.get(key)
    if hit is None:
        K = len(key)
        flat = (ctypes.c_int32 * (3 * K))(*[v for off in key for v in off])
        central = K % 2 == 1 and all(tuple(-v for v in key[K - 1 - k]) == key[k] for k in range(K))
        hit = _offset_cache[key] = (flat, central)
    return hit


def build_kernel_map(in_map: CoordMap, out_map: CoordMap, offsets) -> KernelMap:
    lib = L.load()
    K = len(offsets)
    dev = out_map.coords.device
    flat, central = _offsets_ctypes(offsets)
    nbr = _empty((K, out_map.size), torch.int32, dev)
    tap_count = _empty(K, torch.int32, dev)
    e0 = _profiler.begin() if _profiler else None
    # self map with centrally symmetric offsets (odd kernels at stride 1): half the probes, mirrored writes
    sym = in_map is out_map and K <= 125 and symmetric_maps and central
    if sym:
        L.check(lib.spc_kernel_map_sym(L.ptr(in_map.table), in_map.n_slots, L.ptr(out_map.coords), out_map.size,
                                       ctypes.cast(flat, ctypes.c_void_p), K, L.ptr(nbr), L.ptr(tap_count),
                                       L.stream()), "spc_kernel_map_sym")
    else:
        L.check(lib.spc_kernel_map(L.ptr(in_map.table), in_map.n_slots, L.ptr(out_map.coords), out_map.size,
                                   ctypes.cast(flat, ctypes.c_void_p), K, L.ptr(nbr), L.ptr(tap_count),
                                   L.stream()), "spc_kernel_map")
    if e0 is not None:
        _profiler.end("kernel_map", e0, 0, (16.0 + 8.0 * K + 4.0 * K) * out_map.size,
                      f"K{K} M{out_map.size} ts{in_map.tensor_stride[0]}")
    km = KernelMap(nbr, tap_count, K, in_map.size, out_map.size)
    km.symmetric = bool(sym)
    return km


# ---------------------------------------------------------------------------
# engine-side row order
# ---------------------------------------------------------------------------
# The tensor-core kernels execute a kernel offset for a whole 128-row tile when ANY row of the tile has a neighbour there.
# On the dense synthetic rooms that wastes ~1.45x; on the geometry the reference's ScanNet-plenoxel loader really produces
# (samples 2.3 voxels apart, 5-8 neighbours per voxel at stride 2) a raster-ordered tile touches
# 24-27 offsets for rows that have 5-8 each.  MinkowskiEngine's GPU backend does not define the row order of a map, so
# the engine may choose one: rows with equal neighbour masks are grouped inside windows of `sort_window` rows (windows keep
# the gathers' L2 locality; a global sort loses it — profiles/r2_row_order_experiments.md) when that cuts the executed
# (tile, offset) volume enough to pay for the second map build.  `.C`, `.F`, `unique_index`, `inverse_mapping` and the
# kernel maps of such a map are all in the new order (consistent with each other); results at the points of a
# TensorField (`slice`) do not depend on it.
# OPT-IN: a re-ordered map's exported rows (`.C`, `.F`, kernel-map pair lists) are no longer in MinkowskiEngine's CPU
# (first-occurrence) order, which is the order the bit-exact parity tests pin — so the default keeps that order everywhere.
sort_rows = os.environ.get("SPARSECONV_B200_SORT_ROWS", "0") == "1"
sort_min_rows = 1 << 18      # smaller maps are not worth a second kernel-map build
sort_window = 1 << 16
sort_min_ratio = 2.0         # executed / useful (tile, offset) volume on the first-occurrence order above which rows are re-ordered
sort_stats = {"considered": 0, "reordered": 0}


def reorder_rows_by_mask(cmap: CoordMap, km: KernelMap):
    """Decide from the self map `km` (first-occurrence order) whether grouping rows by neighbour mask pays, and if so
    permute the map in place: coordinates, hash-table rows.  Returns (perm, pos) — new row j' is old row perm[j'], old
    row r is new row pos[r] (int64 / int32 device tensors) — or None when the order is kept.  One host synchronisation."""
    lib = L.load()
    m, K = cmap.size, km.K
    dev = cmap.coords.device
    sort_stats["considered"] += 1
    rmask = _empty(m, torch.int32, dev)
    stat = _empty(1, torch.int64, dev)
    L.check(lib.spc_row_masks(L.ptr(km.nbr), m, K, L.ptr(rmask), L.ptr(stat), L.stream()), "spc_row_masks")
    executed = int(stat.item()) * 128          # (host sync)
    useful = max(km.n_pairs, 1)
    if executed < sort_min_ratio * useful:
        return None
    window = torch.arange(m, device=dev, dtype=torch.int64) // int(sort_window)
    key = (window << 32) | (rmask.to(torch.int64) & 0xFFFFFFFF)
    perm = torch.sort(key, stable=True)[1]
    pos = torch.empty(m, dtype=torch.int32, device=dev)
    pos[perm] = torch.arange(m, device=dev, dtype=torch.int32)
    cmap.coords = cmap.coords[perm].contiguous()
    L.check(lib.spc_table_relabel(L.ptr(cmap.table), cmap.n_slots, L.ptr(pos), L.stream()), "spc_table_relabel")
    sort_stats["reordered"] += 1
    return perm, pos


# ---------------------------------------------------------------------------
# feature-row ops
# ---------------------------------------------------------------------------
def _rows_view(x: torch.Tensor):
    """(tensor, row pitch in elements) for kernels that take a row pitch: a column slice of a wider row-major tensor
    (what torch.cat's backward hands out) is read in place instead of being copied to a dense tensor first."""
    if x.dtype != torch.float32:
        raise RuntimeError(f"features must be float32, got {x.dtype}")
    ensure_filled(x)
    if x.dim() == 2 and x.stride(1) == 1 and x.stride(0) >= x.shape[1] and x.shape[0] > 1:
        return x, x.stride(0)
    x = x.contiguous()
    return x, (x.shape[1] if x.dim() == 2 else 0)


def _ptr_rows(x: torch.Tensor):
    """device pointer of a row-strided CUDA tensor (L.ptr insists on dense tensors)"""
    if not x.is_cuda:
        raise RuntimeError("sparseconv_b200 kernels need CUDA tensors (no CPU fallback)")
    return x.data_ptr()


def _feat(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        raise RuntimeError(f"features must be float32, got {x.dtype}")
    return ensure_filled(x).contiguous()


def segment_reduce(feats, inverse, count, m, mode):
    lib = L.load()
    feats = _feat(feats)
    n, C = feats.shape
    out = _empty((m, C), torch.float32, feats.device)
    L.check(lib.spc_segment_reduce(L.ptr(feats), L.ptr(inverse), L.ptr(count), n, m, C, mode, L.ptr(out),
                                   L.stream()), "spc_segment_reduce")
    return out


def gather_rows(src, index, count=None):
    lib = L.load()
    src = _feat(src)
    n, C = index.shape[0], src.shape[1]
    out = _empty((n, C), torch.float32, src.device)
    L.check(lib.spc_gather_rows(L.ptr(src), L.ptr(index), L.ptr(count), n, C, L.ptr(out), L.stream()),
            "spc_gather_rows")
    return out


def scatter_add_rows(src, index, m):
    lib = L.load()
    src = _feat(src)
    n, C = src.shape
    out = _empty((m, C), torch.float32, src.device)
    L.check(lib.spc_scatter_add_rows(L.ptr(src), L.ptr(index), n, m, C, L.ptr(out), L.stream()),
            "spc_scatter_add_rows")
    return out


class SegmentReduceFn(torch.autograd.Function):
    """TensorField -> SparseTensor feature reduction (mode 0 average, 1 sum, 2 first/subsample)."""

    @staticmethod
    def forward(ctx, feats, inverse, count, first, m, mode):
        ctx.mode = mode
        ctx.n = feats.shape[0]
        if mode == 2:
            ctx.save_for_backward(first)
            return gather_rows(feats, first)
        ctx.save_for_backward(inverse, count)
        return segment_reduce(feats, inverse, count, m, mode)

    @staticmethod
    def backward(ctx, g):
        g = _feat(g)
        if ctx.mode == 2:
            (first,) = ctx.saved_tensors
            return scatter_add_rows(g, first, ctx.n), None, None, None, None, None
        inverse, count = ctx.saved_tensors
        return gather_rows(g, inverse, count if ctx.mode == 0 else None), None, None, None, None, None


class GatherRowsFn(torch.autograd.Function):
    """out[j] = src[index[j]]  (SparseTensor.slice, res16unet.py:435)."""

    @staticmethod
    def forward(ctx, src, index):
        ctx.save_for_backward(index)
        ctx.m = src.shape[0]
        return gather_rows(src, index)

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        return scatter_add_rows(_feat(g), index, ctx.m), None


# ---------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------
_ws_bytes_cache: dict = {}


def _conv_ws_bytes(lib, K, c_in, c_out, precision) -> int:
    key = (K, c_in, c_out, precision)
    v = _ws_bytes_cache.get(key)
    if v is None:
        v = _ws_bytes_cache[key] = int(lib.spc_conv_workspace(K, c_in, c_out, precision))
    return v


def _conv_bytes(km, K, c_in, c_out) -> float:
    """Algorithmic bytes of one conv pass: every feature row once, weights once, dense map once."""
    return 4.0 * km.m_in * c_in + 4.0 * km.m_out * c_out + 4.0 * K * c_in * c_out + 4.0 * K * km.m_out


# bf16 copies written as a side output by the kernel that produced the fp32 rows (BatchNorm apply /
# backward): data_ptr -> (weakref to the fp32 tensor, its version, bf16 tensor).  A convolution in
# bf16 mode takes the copy instead of running a conversion pass over the rows.
_bf16_side: dict = {}


def _remember_bf16(t: torch.Tensor, tb: torch.Tensor) -> None:
    key = t.data_ptr()

    def _drop(ref, key=key):
        e = _bf16_side.get(key)
        if e is not None and e[0] is ref:
            del _bf16_side[key]
    _bf16_side[key] = (weakref.ref(t, _drop), t._version, tb)


def _lookup_bf16(t: torch.Tensor):
    e = _bf16_side.get(t.data_ptr())
    if e is None:
        return None
    ref, version, tb = e
    o = ref()
    if o is None or o.shape != t.shape or tb.shape != t.shape or t._version != version or not t.is_contiguous():
        return None
    return tb


def _want_bf16_side(m: int, C: int) -> bool:
    return _default_precision == L.PREC_BF16 and C % 32 == 0 and C <= 512 and m > 0


def to_bf16(x: torch.Tensor, pad_to: int = 0) -> torch.Tensor:
    """fp32 rows -> dense bf16 copy (round to nearest even) for the SPC_PREC_BF16 kernels, optionally zero-padded
    to `pad_to` columns; a column slice of a wider tensor is converted in place (no dense fp32 copy first)."""
    lib = L.load()
    side = _lookup_bf16(x) if not pad_to or pad_to == x.shape[1] else None
    if side is not None:
        return side
    if x.dim() != 2:
        raise RuntimeError("to_bf16 expects [rows, C] features")
    x, pitch = _rows_view(x)
    m, C = x.shape
    c_dst = max(C, int(pad_to))
    out = torch.empty((m, c_dst), dtype=torch.bfloat16, device=x.device)
    e0 = _profiler.begin() if _profiler else None
    L.check(lib.spc_to_bf16(_ptr_rows(x), m, C, pitch, c_dst, L.ptr(out), L.stream()), "spc_to_bf16")
    if e0 is not None:
        _profiler.end("to_bf16", e0, 0, 4.0 * m * C + 2.0 * m * c_dst)
    return out


# ---------------------------------------------------------------------------------------------------------
# Packed weights.  The tensor-core kernels read weights as pre-swizzled shared-memory slabs
# (spc_conv_pack_weights).  The image does not depend on the launch, so a layer packs ONCE per optimiser step per
# direction (forward: W, dgrad: W^T) instead of once per launch: the cache is keyed on the weight tensor OBJECT (a
# weak reference: a different tensor that re-uses the address misses), its autograd version counter and
# `_weights_epoch`, which `sgd_step` bumps because the fused optimiser kernel writes parameters behind autograd's back.
# ---------------------------------------------------------------------------------------------------------
batch_repack = True    # sgd_step re-packs all cached weight images in one launch (off: each layer packs on first use)
_pack_cache: dict = {}
_weights_epoch = 0
pack_stats = {"hits": 0, "misses": 0}


def invalidate_packed_weights() -> None:
    """Call after changing parameters through a raw pointer (anything autograd's version counter does not see)."""
    global _weights_epoch
    _weights_epoch += 1


def _packed_weights(w3: torch.Tensor, dgrad: int, precision: int, owner: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`dgrad`: 0 forward image, 1 W^T, 2 W^T with reversed offsets (symmetric self maps).
    `owner`: the long-lived tensor (the layer's Parameter) `w3` is a view of — identity and version are taken
    from it (a fresh view object per call, or the tensor autograd hands back in backward, would never hit)."""
    lib = L.load()
    K, c_in, c_out = w3.shape
    obj = owner if owner is not None else w3
    key = (w3.data_ptr(), int(dgrad), precision, K, c_in, c_out)
    ver = (obj._version, _weights_epoch)
    e = _pack_cache.get(key)
    if e is not None and e[0]() is obj and e[1] == ver:
        pack_stats["hits"] += 1
        return e[2]
    pack_stats["misses"] += 1
    nbytes = int(lib.spc_conv_packed_bytes(K, c_in, c_out))
    buf = e[2] if (e is not None and e[2].numel() >= nbytes + 1024) else torch.empty(nbytes + 1024, dtype=torch.uint8,
                                                                                       device=w3.device)
    off = (-buf.data_ptr()) % 1024
    L.check(lib.spc_conv_pack_weights(L.ptr(w3), K, c_in, c_out, int(dgrad), precision, buf.data_ptr() + off,
                                      L.stream()), "spc_conv_pack_weights")

    def _drop(ref, key=key):
        cur = _pack_cache.get(key)
        if cur is not None and cur[0] is ref:
            del _pack_cache[key]
    _pack_cache[key] = (weakref.ref(obj, _drop), ver, buf)
    return buf


_pack_desc = {"sig": None, "table": None}


def repack_all() -> int:
    """Re-pack every cached weight image whose owner (a live Parameter) is stale, in ONE launch
    (spc_conv_pack_weights_batch) — `sgd_step` calls this right after the fused optimiser kernel, so the ~2 x 64
    per-layer pack launches of a Res16UNet34C step become one.  Returns the number of images packed."""
    lib = L.load()
    rows, keys = [], []
    for key, (ref, ver, buf) in list(_pack_cache.items()):
        obj = ref()
        ptr, dgrad, precision, K, c_in, c_out = key
        if obj is None or not obj.is_cuda or obj.data_ptr() != ptr or obj.numel() != K * c_in * c_out \
                or not obj.is_contiguous() or obj.device.index != torch.cuda.current_device():
            continue   # temporaries / moved parameters: packed on demand
        ck, cn = (c_out, c_in) if dgrad else (c_in, c_out)
        rows.append((ptr, _packed_ptr(buf), K, ck, cn, (0, 1, 3)[int(dgrad)], int(precision == L.PREC_BF16), 0))
        keys.append(key)
    if not rows:
        return 0
    sig = tuple(rows)
    if _pack_desc["sig"] != sig:
        _pack_desc["sig"] = sig
        _pack_desc["table"] = torch.tensor(rows, dtype=torch.int64).to(_pack_cache[keys[0]][2].device)
    L.check(lib.spc_conv_pack_weights_batch(L.ptr(_pack_desc["table"]), len(rows), L.stream()),
            "spc_conv_pack_weights_batch")
    for key in keys:
        ref, _, buf = _pack_cache[key]
        _pack_cache[key] = (ref, (ref()._version, _weights_epoch), buf)
    return len(rows)


def _packed_ptr(buf: torch.Tensor) -> int:
    return buf.data_ptr() + ((-buf.data_ptr()) % 1024)


_tc_cache: dict = {}


def _tensor_core(what: int, K: int, c_in: int, c_out: int, precision: int) -> bool:
    key = (what, K, c_in, c_out, precision)
    v = _tc_cache.get(key)
    if v is None:
        v = _tc_cache[key] = bool(L.load().spc_conv_tensor_core(what, K, c_in, c_out, precision))
    return v


# ---------------------------------------------------------------------------------------------------------
# Residual gradients.  In a residual block the input rows x feed conv1 AND the block's `out += residual`
# (resnet_block.py:53-69), so autograd adds two gradients of x with a separate pass over both (26 such passes per
# Res16UNet34C step).  Here a convolution node (SparseConvFn / ConvBNFn) hands its input back as a second output — an
# ALIAS the ME surface swaps into the input's SparseTensor — so the residual branch's gradient arrives at the SAME
# node as `g_alias`; when that gradient is a buffer one of our backward kernels has just written for exactly this
# tensor (`_aim_grad`: BatchNorm backward's dres, a convolution's dx), dgrad reduce-adds into it in its epilogue
# (spc_conv_dgrad_packed_acc) and returns it.  Anything else (a sum formed by autograd, a strided slice of a
# concatenation's gradient) is added with one ordinary pass.
# ---------------------------------------------------------------------------------------------------------
fuse_residual_grad = True
fuse_bn_stats = True         # ConvBNFn: BatchNorm statistics from the convolution's epilogue (no statistics pass over the rows)
_grad_aims: dict = {}    # data_ptr of a gradient buffer -> (weak reference to it, data_ptr of the tensor it is the gradient of)
residual_stats = {"accumulated": 0, "added": 0}


def _aim_grad(g: torch.Tensor, target_ptr: int) -> None:
    key = g.data_ptr()

    def _drop(ref, key=key, table=_grad_aims):
        e = table.get(key)
        if e is not None and e[0] is ref:
            del table[key]
    _grad_aims[key] = (weakref.ref(g, _drop), target_ptr)


def _aimed_at(g: torch.Tensor, target_ptr: int) -> bool:
    """True (once) if `g` is a dense fp32 buffer written by one of our backward kernels as the gradient of the tensor
    at `target_ptr` — nobody else holds it, so it may be accumulated into."""
    e = _grad_aims.get(g.data_ptr())
    if e is None or e[0]() is not g or e[1] != target_ptr:
        return False
    del _grad_aims[g.data_ptr()]
    return g.dtype == torch.float32 and g.is_contiguous() and g.dim() == 2


def _finish_dx(dx_fn, g_alias, x_ptr, shape, tc: bool):
    """Input gradient of a convolution node whose input alias received `g_alias` (or None).  `dx_fn(add_into)` runs
    dgrad, into `add_into` when given."""
    if g_alias is None:
        dx = dx_fn(None)
    elif tc and tuple(g_alias.shape) == tuple(shape) and _aimed_at(g_alias, x_ptr):
        dx = dx_fn(g_alias)
        residual_stats["accumulated"] += 1
    else:
        dx = dx_fn(None)
        dx.add_(ensure_filled(g_alias))
        residual_stats["added"] += 1
    _aim_grad(dx, x_ptr)
    return dx


# ---------------------------------------------------------------------------------------------------------
# Gradient sinks.  A trainer that keeps every parameter gradient in a flat arena (zeroed once per step) registers its
# parameters here; the backward kernels then ADD a parameter's gradient straight into its arena slice (wgrad:
# spc_conv_wgrad_acc; BatchNorm: dgamma / dbeta written by the reduction's last block) and the autograd Function
# returns None for it, instead of materialising a temporary that autograd adds to `.grad` with one tiny launch per
# parameter (267 of them per Res16UNet34C step).  `on_ready(param)` is called once the kernel is enqueued — the
# trainer's bucketed all-reduce hook, which autograd would otherwise fire after its accumulation.
# ---------------------------------------------------------------------------------------------------------
_grad_sinks: dict = {}   # id(param) -> (weak reference to the parameter, on_ready); tensors compare elementwise


def register_grad_sink(param: torch.Tensor, on_ready) -> None:
    key = id(param)

    def _drop(ref, key=key):
        e = _grad_sinks.get(key)
        if e is not None and e[0] is ref:
            del _grad_sinks[key]
    _grad_sinks[key] = (weakref.ref(param, _drop), on_ready)


def clear_grad_sinks() -> None:
    _grad_sinks.clear()


def _sink(param):
    """(grad tensor to add into, on_ready) if `param` is registered and its .grad is a usable dense fp32 buffer."""
    if param is None or not _grad_sinks:
        return None
    e = _grad_sinks.get(id(param))
    if e is None or e[0]() is not param:
        return None
    cb = e[1]
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda or not param.requires_grad:
        return None
    return g, cb


def _drop_offsets(w, bits):
    """CUDA-core kernels take no offset mask: zero the kernels of the dropped offsets instead."""
    keep = torch.tensor([(int(bits) >> k) & 1 for k in range(w.shape[0])], dtype=w.dtype, device=w.device)
    return w * keep.view(-1, 1, 1)


def conv_fwd_raw(x, w, bias, km: KernelMap, precision, w_owner=None, offset_bits: Optional[int] = None,
                 bn_sums: Optional[torch.Tensor] = None):
    """`bn_sums` (float64 [2 * c_out], tensor-core shapes): ask the kernel's epilogue for the per-column sum / sum of
    squares of the output; returns (out, fused) then — fused False: the sums were not produced (small map, wide or
    biased layer) and the caller runs the statistics pass."""
    lib = L.load()
    K, c_in, c_out = w.shape
    out = _empty((km.m_out, c_out), torch.float32, x.device)
    mask = km.masked(offset_bits) if precision != L.PREC_FP32 else None
    if offset_bits is not None and not _tensor_core(0, K, c_in, c_out, precision):
        w, w_owner = _drop_offsets(w, offset_bits), None
    e0 = _profiler.begin() if _profiler else None
    if _tensor_core(0, K, c_in, c_out, precision):
        wp = _packed_weights(w, 0, precision, w_owner)
        if bn_sums is not None:
            fused = ctypes.c_int32(0)
            L.check(lib.spc_conv_fwd_packed_stats(L.ptr(x), _packed_ptr(wp), L.ptr(bias), L.ptr(km.nbr), L.ptr(mask),
                                                  km.m_in, km.m_out, c_in, c_out, K, precision, L.ptr(out),
                                                  L.ptr(bn_sums), ctypes.byref(fused), L.stream()),
                    "spc_conv_fwd_packed_stats")
            if e0 is not None:
                _profiler.end("conv_fwd", e0, 2.0 * km.n_pairs * c_in * c_out, _conv_bytes(km, K, c_in, c_out),
                              f"K{K} {c_in}->{c_out} M{km.m_out} P{km.n_pairs}")
            return out, bool(fused.value)
        L.check(lib.spc_conv_fwd_packed(L.ptr(x), _packed_ptr(wp), L.ptr(bias), L.ptr(km.nbr), L.ptr(mask), km.m_in,
                                        km.m_out, c_in, c_out, K, precision, L.ptr(out), L.stream()),
                "spc_conv_fwd_packed")
    else:
        ws_bytes = _conv_ws_bytes(lib, K, c_in, c_out, precision)
        ws = _workspace(ws_bytes, x.device)
        L.check(lib.spc_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(km.nbr), L.ptr(mask), km.m_in, km.m_out,
                                 c_in, c_out, K, precision, L.ptr(out), L.ptr(ws), ws_bytes, L.stream()),
                "spc_conv_fwd")
    if e0 is not None:
        _profiler.end("conv_fwd", e0, 2.0 * km.n_pairs * c_in * c_out, _conv_bytes(km, K, c_in, c_out),
                      f"K{K} {c_in}->{c_out} M{km.m_out} P{km.n_pairs}")
    return out


def conv_dgrad_raw(g, w, km: KernelMap, precision, w_owner=None, offset_bits: Optional[int] = None,
                   add_into: Optional[torch.Tensor] = None):
    """`add_into` (tensor-core shapes only): a dense fp32 [m_in, c_in] buffer the input gradient is ADDED to."""
    lib = L.load()
    K, c_in, c_out = w.shape
    din = add_into if add_into is not None else _empty((km.m_in, c_in), torch.float32, g.device)
    tc = _tensor_core(1, K, c_in, c_out, precision)
    # symmetric self map: the forward map with reversed offsets IS the transposed map (no transpose, no second mask)
    sym = tc and km.symmetric and offset_bits is None and symmetric_dgrad
    if offset_bits is not None and not tc:
        w, w_owner = _drop_offsets(w, offset_bits), None
    if sym:
        nbr_t, mask_t = km.nbr, km.mask
    else:
        nbr_t = km.nbr_t
        mask_t = km.masked(offset_bits, transposed=True) if precision != L.PREC_FP32 else None
    e0 = _profiler.begin() if _profiler else None
    if tc:
        wp = _packed_weights(w, 2 if sym else 1, precision, w_owner)
        L.check(lib.spc_conv_dgrad_packed_acc(L.ptr(g), _packed_ptr(wp), L.ptr(nbr_t), L.ptr(mask_t), km.m_in, km.m_out,
                                              c_in, c_out, K, precision, L.ptr(din), int(add_into is not None),
                                              L.stream()), "spc_conv_dgrad_packed")
    else:
        if add_into is not None:
            raise RuntimeError("conv_dgrad_raw(add_into=...) needs a tensor-core shape")
        ws_bytes = _conv_ws_bytes(lib, K, c_in, c_out, precision)
        ws = _workspace(ws_bytes, g.device)
        L.check(lib.spc_conv_dgrad(L.ptr(g), L.ptr(w), L.ptr(nbr_t), L.ptr(mask_t), km.m_in, km.m_out, c_in,
                                   c_out, K, precision, L.ptr(din), L.ptr(ws), ws_bytes, L.stream()),
                "spc_conv_dgrad")
    if e0 is not None:
        _profiler.end("conv_dgrad", e0, 2.0 * km.n_pairs * c_in * c_out, _conv_bytes(km, K, c_in, c_out),
                      f"K{K} {c_in}->{c_out} M{km.m_out} P{km.n_pairs}")
    return din


def conv_wgrad_raw(x, g, km: KernelMap, K, c_in, c_out, precision, add_into: Optional[torch.Tensor] = None,
                   offset_bits: Optional[int] = None):
    """dW [K, c_in, c_out]; `add_into` (a dense fp32 buffer of that many elements, tensor-core shapes only): the
    gradient is ADDED to it (spc_conv_wgrad_acc) and it is returned."""
    lib = L.load()
    mask = km.masked(offset_bits) if precision != L.PREC_FP32 else None
    e0 = _profiler.begin() if _profiler else None
    if add_into is not None:
        dw = add_into
        L.check(lib.spc_conv_wgrad_acc(L.ptr(x), L.ptr(g), L.ptr(km.nbr), L.ptr(mask), km.m_in, km.m_out, c_in, c_out,
                                       K, precision, L.ptr(dw), 1, L.stream()), "spc_conv_wgrad_acc")
    else:
        dw = _empty((K, c_in, c_out), torch.float32, x.device)
        L.check(lib.spc_conv_wgrad_acc(L.ptr(x), L.ptr(g), L.ptr(km.nbr), L.ptr(mask), km.m_in, km.m_out, c_in, c_out,
                                       K, precision, L.ptr(dw), 0, L.stream()), "spc_conv_wgrad")
    if e0 is not None:
        _profiler.end("conv_wgrad", e0, 2.0 * km.n_pairs * c_in * c_out, _conv_bytes(km, K, c_in, c_out),
                      f"K{K} {c_in}->{c_out} M{km.m_out} P{km.n_pairs}")
    return dw


class SparseConvFn(torch.autograd.Function):
    """out[o] = sum_k x[nbr[k,o]] @ W[k] (+bias); backward = dgrad / wgrad kernels."""

    @staticmethod
    def forward(ctx, x, w, bias, km, precision, w_param=None, offset_bits=None, alias=False):
        """`w_param`: the Parameter `w` is (a view of), for the gradient sink (see register_grad_sink).
        `offset_bits`: bit k set = kernel offset k takes part (weight-sparse inference convolution); None = all.
        `alias`: also return the input rows as a second output (see "Residual gradients")."""
        x_in = x
        w3 = w.contiguous()
        if x.dim() != 2 or x.shape[0] != km.m_in or x.shape[1] != w3.shape[1]:
            raise RuntimeError(f"conv input {tuple(x.shape)} does not match map rows {km.m_in} / kernel {tuple(w3.shape)}")
        b = bias.contiguous().view(-1) if bias is not None else None
        K, c_in, c_out = w3.shape
        # The tensor-core kernels take channel counts that are multiples of 32; on large maps zero-pad
        # odd widths (27 SH channels, 20 classes) instead of dropping to the CUDA-core kernels.
        pad_in = pad_out = 0
        big = max(km.m_in, km.m_out) >= 4096
        if precision == L.PREC_BF16 and (K > 32 or c_out > 1024 or K * (c_in + (-c_in) % 32) > 128 * 128
                                         or (not big and (c_in % 32 or c_out % 32))):
            precision = L.PREC_TF32  # shapes the bf16 kernels are not built for (bench.py reports the routes taken)
        if precision != L.PREC_FP32 and K <= 32 and big:
            pad_in, pad_out = (-c_in) % 32, (-c_out) % 32
        if precision != L.PREC_BF16:
            x = _feat(x)   # (bf16 mode reads the operand copy: a hollow input is not filled for it)
        if pad_in and precision != L.PREC_BF16:
            x = torch.nn.functional.pad(x, (0, pad_in))
        if pad_in or pad_out:
            w3 = torch.nn.functional.pad(w3, (0, pad_out, 0, pad_in))
            if b is not None and pad_out:
                b = torch.nn.functional.pad(b, (0, pad_out))
        if precision == L.PREC_BF16:
            # conversion and channel padding in one pass; the bf16 copy is what backward needs too
            x = to_bf16(x, pad_to=c_in + pad_in)
        own = w_param if (w_param is not None and not pad_in and not pad_out
                          and w_param.data_ptr() == w3.data_ptr() and w_param.numel() == w3.numel()) else None
        out = conv_fwd_raw(x, w3, b, km, precision, own, offset_bits)
        ctx.offset_bits = offset_bits
        if pad_out:
            out = out[:, :c_out].contiguous()
        ctx.save_for_backward(x, w3)
        ctx.w_param = own
        ctx.km = km
        ctx.precision = precision
        ctx.has_bias = bias is not None
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.dims = (c_in, c_out, pad_in, pad_out)
        ctx.x_ptr = x_in.data_ptr()
        ctx.alias = bool(alias)
        if alias:
            ctx.set_materialize_grads(False)   # an output nobody used arrives as None in backward, not as zeros
            return out, x_in.view_as(x_in)
        return out

    @staticmethod
    def backward(ctx, g, g_alias=None):
        x, w3 = ctx.saved_tensors
        km, prec = ctx.km, ctx.precision
        c_in, c_out, pad_in, pad_out = ctx.dims
        dx = dw = db = None
        if g is None:   # only the alias was used downstream
            return (g_alias if ctx.needs_input_grad[0] else None), None, None, None, None, None, None, None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = g.sum(0).view(ctx.bias_shape)
        if prec == L.PREC_BF16:
            g = to_bf16(g, pad_to=c_out + pad_out)  # converts, de-strides and pads in one pass
        else:
            g = _feat(g)
            if pad_out:
                g = torch.nn.functional.pad(g, (0, pad_out))
        if ctx.needs_input_grad[0]:
            def run_dgrad(add_into):
                d = conv_dgrad_raw(g, w3, km, prec, ctx.w_param, ctx.offset_bits, add_into=add_into)
                return d[:, :c_in].contiguous() if pad_in else d
            tc = (not pad_in) and ctx.offset_bits is None and _tensor_core(1, *w3.shape, prec)
            dx = _finish_dx(run_dgrad, g_alias, ctx.x_ptr, (km.m_in, c_in), tc)
        if ctx.needs_input_grad[1]:
            K, ci, co = w3.shape
            sink = _sink(ctx.w_param) if (_tensor_core(2, K, ci, co, prec) and ctx.offset_bits is None) else None
            if sink is not None and sink[0].numel() == K * ci * co:
                conv_wgrad_raw(x, g, km, K, ci, co, prec, add_into=sink[0])   # straight into the gradient arena
                sink[1](ctx.w_param)
            else:
                dw = conv_wgrad_raw(x, g, km, K, ci, co, prec, offset_bits=ctx.offset_bits)
                if ctx.offset_bits is not None and not _tensor_core(2, K, ci, co, prec):
                    dw = _drop_offsets(dw, ctx.offset_bits)
                if pad_in or pad_out:
                    dw = dw[:, :c_in, :c_out].contiguous()
        return dx, dw, db, None, None, None, None, None


def conv_bn_fusable(x, w, bias, km: KernelMap, precision: int, offset_bits) -> bool:
    """True when convolution + BatchNorm can run as the one autograd node `ConvBNFn`: bf16 operands, tensor-core
    shapes in all three directions without channel padding, no bias, all offsets."""
    if not fuse_conv_bn or precision != L.PREC_BF16 or _default_precision != L.PREC_BF16:
        return False
    if bias is not None or offset_bits is not None or w.dim() != 3:
        return False
    K, c_in, c_out = w.shape
    if K > 32 or c_in % 32 or c_out % 32 or c_out > 512 or K * c_in > 128 * 128 or km.m_out < 1:
        return False
    return (_tensor_core(0, K, c_in, c_out, precision) and _tensor_core(1, K, c_in, c_out, precision)
            and _tensor_core(2, K, c_in, c_out, precision))


class ConvBNFn(torch.autograd.Function):
    """MinkowskiConvolution -> MinkowskiBatchNorm (-> += residual) (-> ReLU) as ONE autograd node (bf16 operand mode).

    The convolution output `c` is internal to the node, so backward never materialises its fp32 gradient: BatchNorm
    backward writes the bf16 operand copy only (4 of its 26 bytes per element less) and dgrad / wgrad read that.
    Arithmetic and kernels are those of `SparseConvFn` followed by `BatchNormFn`."""

    @staticmethod
    def forward(ctx, x, w, km, w_param, gamma, beta, running_mean, running_var, training, momentum, eps, relu,
                residual, tracked, want_fp32, alias=False):
        w3 = w.contiguous()
        K, c_in, c_out = w3.shape
        if x.dim() != 2 or x.shape[0] != km.m_in or x.shape[1] != c_in:
            raise RuntimeError(f"conv input {tuple(x.shape)} does not match map rows {km.m_in} / kernel {tuple(w3.shape)}")
        xb = to_bf16(x)
        own = w_param if (w_param is not None and w_param.data_ptr() == w3.data_ptr()
                          and w_param.numel() == w3.numel()) else None
        sums = None
        if fuse_bn_stats and (training or running_mean is None) and c_out <= 128 and km.m_out >= 1:
            sums = _empty(2 * c_out, torch.float64, x.device)
            c, fused = conv_fwd_raw(xb, w3, None, km, L.PREC_BF16, own, bn_sums=sums)
            if not fused:
                sums = None
        else:
            c = conv_fwd_raw(xb, w3, None, km, L.PREC_BF16, own)
        y, yb, mean, var, use_batch = _bn_forward_impl(c, gamma, beta, running_mean, running_var, training, momentum,
                                                       eps, relu, residual, tracked, want_fp32, sums)
        relu_mode = 0 if not relu else (2 if (residual is None and recompute_relu_mask) else 1)
        ctx.save_for_backward(xb, w3, c, ((yb if yb is not None else y) if relu_mode == 1 else None), mean, var,
                              gamma, beta)
        ctx.cfg = (float(eps), relu_mode, int(use_batch), residual is not None)
        ctx.params = (gamma, beta)
        ctx.w_param = own
        ctx.km = km
        ctx.x_ptr = x.data_ptr()
        ctx.res_ptr = residual.data_ptr() if residual is not None else None
        if alias:
            ctx.set_materialize_grads(False)   # an output nobody used arrives as None in backward, not as zeros
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, g_alias=None):
        xb, w3, c, ymask, mean, var, gamma, beta = ctx.saved_tensors
        eps, relu_mode, use_batch, has_res = ctx.cfg
        km = ctx.km
        K, c_in, c_out = w3.shape
        if dy is None:   # only the alias was used downstream
            return (g_alias,) + (None,) * 15
        dc, dcb, dres, dgamma, dbeta = _bn_backward_impl(c, ymask, dy, mean, var, gamma, beta, eps, relu_mode,
                                                         use_batch, has_res, ctx.params, want_dx32=False,
                                                         res_ptr=ctx.res_ptr)
        g = dcb if dcb is not None else to_bf16(dc)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _finish_dx(lambda add_into: conv_dgrad_raw(g, w3, km, L.PREC_BF16, ctx.w_param, add_into=add_into),
                            g_alias, ctx.x_ptr, (km.m_in, c_in), True)
        if ctx.needs_input_grad[1]:
            sink = _sink(ctx.w_param)
            if sink is not None and sink[0].numel() == K * c_in * c_out:
                conv_wgrad_raw(xb, g, km, K, c_in, c_out, L.PREC_BF16, add_into=sink[0])
                sink[1](ctx.w_param)
            else:
                dw = conv_wgrad_raw(xb, g, km, K, c_in, c_out, L.PREC_BF16)
        return dx, dw, None, None, dgamma, dbeta, None, None, None, None, None, None, dres, None, None, None


# ---------------------------------------------------------------------------
# batch norm / relu / add
# ---------------------------------------------------------------------------
# "Hollow" rows.  In bf16 mode a BatchNorm output that only feeds convolutions is needed as the bf16 operand copy
# alone; its fp32 tensor still has to exist (it is the autograd output) but is left UNWRITTEN — 4 of the 14 bytes per
# element the apply pass moves.  `_hollow` maps such a tensor to a closure that writes its fp32 rows on demand
# (`ensure_filled`, called by every reader of fp32 rows in this module and by `Tensor.F` at the ME surface).
_hollow: dict = {}   # data_ptr -> (weak reference to the tensor, fill closure taking the tensor)
hollow_stats = {"made": 0, "filled": 0}


def _mark_hollow(t: torch.Tensor, fill) -> None:
    key = t.data_ptr()

    def _drop(ref, key=key, table=_hollow):
        e = table.get(key)
        if e is not None and e[0] is ref:
            del table[key]
    _hollow[key] = (weakref.ref(t, _drop), fill)
    hollow_stats["made"] += 1


def _hollow_entry(t: torch.Tensor):
    """registry entry of `t` or of the tensor `t` is a whole-tensor alias of (ConvBNFn / SparseConvFn hand their input
    back as a second output, see "Residual gradients")"""
    e = _hollow.get(t.data_ptr()) if _hollow else None
    if e is None:
        return None
    o = e[0]()
    if o is t or (o is not None and t._base is o and t.shape == o.shape and t.is_contiguous()):
        return e
    return None


def is_hollow(t: torch.Tensor) -> bool:
    return _hollow_entry(t) is not None


def ensure_filled(t: torch.Tensor) -> torch.Tensor:
    """Write the fp32 rows of a hollow tensor (no-op for every other tensor)."""
    if _hollow:
        e = _hollow_entry(t)
        if e is not None:
            del _hollow[t.data_ptr()]
            e[1](e[0]())
            hollow_stats["filled"] += 1
    return t


def _bn_forward_impl(x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, residual, tracked,
                     want_fp32=True, sums=None):
    """Statistics + apply.  Returns (y, yb, mean, var, use_batch).  `want_fp32=False`: y may be left hollow.
    `sums` (float64 [2C]): per-column sum / sum of squares already produced by the convolution's epilogue — the
    statistics pass over x is replaced by spc_bn_finalize."""
    lib = L.load()
    m, C = x.shape
    dev = x.device
    ws_bytes = int(lib.spc_bn_workspace(m, C))
    ws = _workspace(ws_bytes, dev)
    use_batch = training or running_mean is None
    e0 = _profiler.begin() if _profiler else None
    if use_batch:
        if m < 1:
            raise RuntimeError("batch norm over zero rows")
        mean = _empty(C, torch.float32, dev)
        var = _empty(C, torch.float32, dev)
        upd = training and running_mean is not None
        if sums is not None:
            L.check(lib.spc_bn_finalize(L.ptr(sums), m, C, L.ptr(mean), L.ptr(var),
                                        L.ptr(running_mean) if upd else None, L.ptr(running_var) if upd else None,
                                        float(momentum), L.ptr(tracked) if (upd and tracked is not None) else None,
                                        L.stream()), "spc_bn_finalize")
        else:
            L.check(lib.spc_bn_stats_tracked(L.ptr(x), m, C, L.ptr(mean), L.ptr(var),
                                             L.ptr(running_mean) if upd else None,
                                             L.ptr(running_var) if upd else None, float(momentum),
                                             L.ptr(tracked) if (upd and tracked is not None) else None, L.ptr(ws),
                                             ws_bytes, L.stream()), "spc_bn_stats")
    else:
        mean, var = running_mean, running_var
    res = _feat(residual) if residual is not None else None
    y = _empty((m, C), torch.float32, dev)
    yb = _empty((m, C), torch.bfloat16, dev) if _want_bf16_side(m, C) else None
    hollow = yb is not None and not want_fp32 and hollow_rows
    L.check(lib.spc_bn_apply(L.ptr(x), L.ptr(mean), L.ptr(var), L.ptr(gamma), L.ptr(beta), L.ptr(res), m, C,
                             float(eps), int(relu), None if hollow else L.ptr(y), L.ptr(yb), L.stream()), "spc_bn_apply")
    if yb is not None:
        _remember_bf16(y, yb)
    if hollow:
        epoch = _weights_epoch

        # (the closure must not hold `y`: the registry entry lives exactly as long as the tensor does)
        def fill(y, x=x, mean=mean, var=var, gamma=gamma, beta=beta, res=res, eps=float(eps), relu=int(relu)):
            if epoch != _weights_epoch:
                raise RuntimeError("fp32 rows of a BatchNorm output were requested after an optimiser step changed its "
                                   "parameters; read .F before stepping (or set ops.hollow_rows = False)")
            L.check(L.load().spc_bn_apply(L.ptr(x), L.ptr(mean), L.ptr(var), L.ptr(gamma), L.ptr(beta), L.ptr(res),
                                          x.shape[0], x.shape[1], eps, relu, L.ptr(y), None, L.stream()),
                    "spc_bn_apply(fill)")
        _mark_hollow(y, fill)
    if e0 is not None:
        _profiler.end("bn_fwd", e0, 0, ((8.0 if (use_batch and sums is None) else 4.0) + (0.0 if hollow else 4.0)
                                        + (4.0 if res is not None else 0.0)
                                        + (2.0 if yb is not None else 0.0)) * m * C, f"C{C} M{m}")
    return y, yb, mean, var, use_batch


def _bn_backward_impl(x, ymask, dy, mean, var, gamma, beta, eps, relu_mode, use_batch, has_res, params,
                      want_dx32=True, res_ptr=None):
    """BatchNorm backward.  `relu_mode`: 0 none, 1 mask from `ymask` (fp32 rows or their bf16 copy), 2 mask
    re-computed from x (no residual before the ReLU: nothing but x and dy is read).  Returns (dx, dxb, dres, dgamma,
    dbeta); dx is None when `want_dx32` is False and a bf16 copy is produced (fused conv + BN node)."""
    lib = L.load()
    dy, dy_pitch = _rows_view(dy)
    m, C = x.shape
    dev = x.device
    ws_bytes = int(lib.spc_bn_workspace(m, C))
    ws = _workspace(ws_bytes, dev)
    dxb = _empty((m, C), torch.bfloat16, dev) if _want_bf16_side(m, C) else None
    dx = _empty((m, C), torch.float32, dev) if (want_dx32 or dxb is None) else None
    dres = _empty((m, C), torch.float32, dev) if has_res else None
    # parameter gradients straight into the trainer's arena when both parameters are registered sinks
    gp, bp = params
    affine = gamma is not None
    sg, sb = (_sink(gp), _sink(bp)) if affine else (None, None)
    direct = sg is not None and sb is not None and sg[0].numel() == C and sb[0].numel() == C
    if direct:
        dgamma, dbeta = sg[0], sb[0]
    else:
        dgamma = _empty(C, torch.float32, dev)
        dbeta = _empty(C, torch.float32, dev)
    e0 = _profiler.begin() if _profiler else None
    y32, y16 = (None, ymask) if (ymask is not None and ymask.dtype == torch.bfloat16) else (ymask, None)
    L.check(lib.spc_bn_bwd_acc(L.ptr(x), L.ptr(y32), L.ptr(y16), _ptr_rows(dy), dy_pitch, L.ptr(mean), L.ptr(var),
                               L.ptr(gamma), L.ptr(beta), m, C, eps,
                               relu_mode, use_batch, L.ptr(dx), L.ptr(dxb), L.ptr(dres), L.ptr(dgamma), L.ptr(dbeta),
                               int(direct), L.ptr(ws), ws_bytes, L.stream()), "spc_bn_bwd")
    if direct:
        sg[1](gp)
        sb[1](bp)
        dgamma = dbeta = None
    if dxb is not None and dx is not None:
        _remember_bf16(dx, dxb)
    if dres is not None and res_ptr is not None:
        _aim_grad(dres, res_ptr)   # (a convolution node that handed out the residual rows may accumulate into it)
    if e0 is not None:
        _profiler.end("bn_bwd", e0, 0, (16.0 + ((4.0 if y16 is not None else 8.0) if relu_mode == 1 else 0.0)
                                        + (4.0 if dx is not None else 0.0) + (4.0 if has_res else 0.0)
                                        + (2.0 if dxb is not None else 0.0)) * m * C,
                      f"C{C} M{m}")
    return dx, dxb, dres, (dgamma if affine else None), (dbeta if affine else None)


class BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm1d semantics on [M,C] rows, optional fused ReLU and residual add."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, residual,
                tracked=None, want_fp32=True):
        """`tracked`: nn.BatchNorm1d.num_batches_tracked (int64 device scalar) to increment inside the statistics
        kernel, or None.  `want_fp32=False`: the caller only needs the bf16 operand copy of the output (see
        "Hollow rows" above)."""
        x = _feat(x)
        y, yb, mean, var, use_batch = _bn_forward_impl(x, gamma, beta, running_mean, running_var, training, momentum,
                                                       eps, relu, residual, tracked, want_fp32)
        # ReLU mask of backward: re-computed from x when nothing was added before the ReLU (mode 2); otherwise from
        # the bf16 copy of y when there is one (2 instead of 4 bytes per value)
        relu_mode = 0 if not relu else (2 if (residual is None and recompute_relu_mask) else 1)
        ctx.save_for_backward(x, ((yb if yb is not None else y) if relu_mode == 1 else None), mean, var, gamma, beta)
        ctx.cfg = (float(eps), relu_mode, int(use_batch), residual is not None)
        ctx.params = (gamma, beta)   # the Parameter objects, for the gradient sinks
        ctx.res_ptr = residual.data_ptr() if residual is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, ymask, mean, var, gamma, beta = ctx.saved_tensors
        eps, relu_mode, use_batch, has_res = ctx.cfg
        dx, _, dres, dgamma, dbeta = _bn_backward_impl(x, ymask, dy, mean, var, gamma, beta, eps, relu_mode, use_batch,
                                                       has_res, ctx.params, res_ptr=ctx.res_ptr)
        return dx, dgamma, dbeta, None, None, None, None, None, None, dres, None, None


def cat_rows_bf16(parts) -> torch.Tensor:
    """bf16 operand copy of the channel concatenation of `parts` ([M, C_i] fp32 rows): each part's bf16 copy (the side
    output of the kernel that made it, else a conversion pass) is copied into its column slice — 4 bytes per element
    moved instead of the 8 of an fp32 torch.cat plus the 6 of its conversion."""
    lib = L.load()
    m = parts[0].shape[0]
    C = sum(int(p.shape[1]) for p in parts)
    out = torch.empty((m, C), dtype=torch.bfloat16, device=parts[0].device)
    e0 = _profiler.begin() if _profiler else None
    c0 = 0
    for p in parts:
        pb = to_bf16(p)
        ci = int(p.shape[1])
        L.check(lib.spc_copy_rows(L.ptr(pb), 2 * ci, out.data_ptr() + 2 * c0, 2 * C, 2 * ci, m, L.stream()),
                "spc_copy_rows")
        c0 += ci
    if e0 is not None:
        _profiler.end("cat_bf16", e0, 0, 4.0 * m * C, f"C{C} M{m}")
    return out


class CatFn(torch.autograd.Function):
    """ME.cat along the channels whose consumers are bf16 convolutions: the fp32 result is HOLLOW (allocated for
    autograd, written only if somebody asks for `.F`), the bf16 operand copy is assembled from the parts' copies."""

    @staticmethod
    def forward(ctx, *parts):
        m = parts[0].shape[0]
        widths = [int(p.shape[1]) for p in parts]
        y = _empty((m, sum(widths)), torch.float32, parts[0].device)
        yb = cat_rows_bf16(parts)
        _remember_bf16(y, yb)

        def fill(y, parts=parts):
            torch.cat([ensure_filled(p) for p in parts], dim=1, out=y)
        _mark_hollow(y, fill)
        ctx.widths = widths
        return y

    @staticmethod
    def backward(ctx, g):
        out, c0 = [], 0
        for w in ctx.widths:
            out.append(g[:, c0:c0 + w])   # column slices, read in place by the consumers (row pitch)
            c0 += w
        return tuple(out)


class SyncBatchNormFn(torch.autograd.Function):
    """BatchNorm whose statistics span the rows of ALL ranks (ME.MinkowskiSyncBatchNorm, train.py:106-107).

    forward: local per-channel statistics from `spc_bn_stats`, turned into (count, sum, sum of squares) and summed
    over the ranks with ONE all-reduce of 2C + 1 doubles, then `spc_bn_apply` with the global mean /
    biased variance; running statistics use the global count (unbiased variance), as torch.nn.SyncBatchNorm.
    backward: the per-channel sums of dy and dy * xhat are summed over the ranks the same way; the weight / bias
    gradients stay LOCAL sums (the data-parallel step averages parameter gradients afterwards)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, group):
        import torch.distributed as dist
        lib = L.load()
        x = _feat(x)
        m, C = x.shape
        dev = x.device
        ws_bytes = int(lib.spc_bn_workspace(max(m, 1), C))
        ws = _workspace(ws_bytes, dev)
        mean = torch.zeros(C, dtype=torch.float32, device=dev)
        var = torch.zeros(C, dtype=torch.float32, device=dev)
        if m > 0:
            L.check(lib.spc_bn_stats(L.ptr(x), m, C, L.ptr(mean), L.ptr(var), None, None, 0.0, L.ptr(ws), ws_bytes,
                                     L.stream()), "spc_bn_stats")
        mu = mean.double()
        packed = torch.cat([torch.tensor([float(m)], dtype=torch.float64, device=dev), mu * m, (var.double() + mu * mu) * m])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        total = packed[0]
        gmean = packed[1:1 + C] / total
        gvar = (packed[1 + C:] / total - gmean * gmean).clamp_min(0)
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(momentum * gmean.float())
                running_var.mul_(1 - momentum).add_(momentum * (gvar * total / (total - 1).clamp_min(1)).float())
        gmean32, gvar32 = gmean.float().contiguous(), gvar.float().contiguous()
        y = _empty((m, C), torch.float32, dev)
        if m > 0:
            L.check(lib.spc_bn_apply(L.ptr(x), L.ptr(gmean32), L.ptr(gvar32), L.ptr(gamma), L.ptr(beta), None, m, C,
                                     float(eps), 0, L.ptr(y), None, L.stream()), "spc_bn_apply")
        ctx.save_for_backward(x, gmean32, gvar32, gamma)
        ctx.cfg = (float(eps), group, total)
        return y

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist
        x, gmean, gvar, gamma = ctx.saved_tensors
        eps, group, total = ctx.cfg
        C = x.shape[1]
        rstd = torch.rsqrt(gvar + eps)
        xhat = (x - gmean) * rstd
        dbeta = dy.sum(0)
        dgamma = (dy * xhat).sum(0)
        packed = torch.cat([dbeta.double(), dgamma.double()])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        s0, s1 = (packed[:C] / total).float(), (packed[C:] / total).float()
        scale = rstd if gamma is None else rstd * gamma
        dx = scale * (dy - s0 - xhat * s1)
        return dx, (dgamma if gamma is not None else None), (dbeta if gamma is not None else None), None, None, None, None, None


class ReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = L.load()
        x = _feat(x)
        y = torch.empty_like(x)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_relu_fwd(L.ptr(x), x.numel(), L.ptr(y), L.stream()), "spc_relu_fwd")
        if e0 is not None:
            _profiler.end("relu_fwd", e0, 0, 8.0 * x.numel())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        (y,) = ctx.saved_tensors
        g = _feat(g)
        dx = torch.empty_like(g)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_relu_bwd(L.ptr(y), L.ptr(g), g.numel(), L.ptr(dx), L.stream()), "spc_relu_bwd")
        if e0 is not None:
            _profiler.end("relu_bwd", e0, 0, 12.0 * g.numel())
        return dx


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = L.load()
        a, b = _feat(a), _feat(b)
        if a.shape != b.shape:
            raise RuntimeError("add: shape mismatch")
        y = torch.empty_like(a)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_add(L.ptr(a), L.ptr(b), a.numel(), L.ptr(y), L.stream()), "spc_add")
        if e0 is not None:
            _profiler.end("add", e0, 0, 12.0 * a.numel())
        return y

    @staticmethod
    def backward(ctx, g):
        return g, g


# ---------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------
class LocalPoolFn(torch.autograd.Function):
    """Sum / average pooling over a kernel map region (kernel_size == stride in the reference)."""

    @staticmethod
    def forward(ctx, x, km, avg):
        lib = L.load()
        x = _feat(x)
        C = x.shape[1]
        out = _empty((km.m_out, C), torch.float32, x.device)
        L.check(lib.spc_pool_fwd(L.ptr(x), L.ptr(km.nbr), km.m_out, C, km.K, int(avg), L.ptr(out), L.stream()),
                "spc_pool_fwd")
        ctx.km = km
        ctx.avg = avg
        return out

    @staticmethod
    def backward(ctx, g):
        # din[i] = sum_k dout[nbr_t[k,i]] (/ count of the out row): a pooling over the swapped map
        lib = L.load()
        km = ctx.km
        g = _feat(g)
        C = g.shape[1]
        if ctx.avg:
            # scale rows of g by 1/(number of present inputs) first
            cnt = (km.nbr >= 0).sum(0).clamp_(min=1).to(torch.float32).unsqueeze(1)
            g = (g / cnt).contiguous()
        din = _empty((km.m_in, C), torch.float32, g.device)
        L.check(lib.spc_pool_fwd(L.ptr(g), L.ptr(km.nbr_t), km.m_in, C, km.K, 0, L.ptr(din), L.stream()),
                "spc_pool_fwd(bwd)")
        return din, None, None


class LocalMaxPoolFn(torch.autograd.Function):
    """Max pooling over a kernel-map region (ME.MinkowskiMaxPooling); backward routes to the arg-max rows."""

    @staticmethod
    def forward(ctx, x, km):
        lib = L.load()
        x = _feat(x)
        C = x.shape[1]
        out = _empty((km.m_out, C), torch.float32, x.device)
        arg = _empty((km.m_out, C), torch.int32, x.device)
        L.check(lib.spc_pool_max_fwd(L.ptr(x), L.ptr(km.nbr), km.m_out, C, km.K, L.ptr(out), L.ptr(arg), L.stream()),
                "spc_pool_max_fwd")
        ctx.save_for_backward(arg)
        ctx.m_in = km.m_in
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        (arg,) = ctx.saved_tensors
        g = _feat(g)
        m_out, C = g.shape
        din = _empty((ctx.m_in, C), torch.float32, g.device)
        L.check(lib.spc_pool_max_bwd(L.ptr(g), L.ptr(arg), m_out, ctx.m_in, C, L.ptr(din), L.stream()),
                "spc_pool_max_bwd")
        return din, None


class GlobalMaxPoolFn(torch.autograd.Function):
    """Per-batch-index maximum over all rows (ME.MinkowskiGlobalMaxPooling)."""

    @staticmethod
    def forward(ctx, x, coords, n_batch):
        lib = L.load()
        x = _feat(x)
        m, C = x.shape
        out = _empty((n_batch, C), torch.float32, x.device)
        arg = _empty((n_batch, C), torch.int32, x.device)
        ws_bytes = n_batch * C * 4
        ws = _workspace(ws_bytes, x.device)
        L.check(lib.spc_global_max_fwd(L.ptr(x), L.ptr(coords), m, C, n_batch, L.ptr(out), L.ptr(arg), L.ptr(ws),
                                       ws_bytes, L.stream()), "spc_global_max_fwd")
        ctx.save_for_backward(arg)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        (arg,) = ctx.saved_tensors
        g = _feat(g)
        nb, C = g.shape
        din = _empty((ctx.m, C), torch.float32, g.device)
        L.check(lib.spc_pool_max_bwd(L.ptr(g), L.ptr(arg), nb, ctx.m, C, L.ptr(din), L.stream()), "spc_pool_max_bwd")
        return din, None, None


class GlobalPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coords, n_batch, avg):
        lib = L.load()
        x = _feat(x)
        m, C = x.shape
        out = _empty((n_batch, C), torch.float32, x.device)
        cnt = _empty(n_batch, torch.int32, x.device)
        L.check(lib.spc_global_pool_fwd(L.ptr(x), L.ptr(coords), m, C, n_batch, int(avg), L.ptr(out), L.ptr(cnt),
                                        L.stream()), "spc_global_pool_fwd")
        ctx.save_for_backward(coords, cnt)
        ctx.cfg = (m, C, n_batch, int(avg))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        coords, cnt = ctx.saved_tensors
        m, C, n_batch, avg = ctx.cfg
        g = _feat(g)
        din = _empty((m, C), torch.float32, g.device)
        L.check(lib.spc_global_pool_bwd(L.ptr(g), L.ptr(coords), L.ptr(cnt), m, C, n_batch, avg, L.ptr(din),
                                        L.stream()), "spc_global_pool_bwd")
        return din, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """mean softmax cross-entropy with ignore_index (nn.CrossEntropyLoss semantics) in one pass."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        lib = L.load()
        logits = _feat(logits)
        if target.dtype != torch.int64:
            raise RuntimeError(f"targets must be int64, got {target.dtype}")
        target = target.contiguous()
        n, C = logits.shape
        if target.shape != (n,):
            raise RuntimeError(f"target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
        graw = _empty((n, C), torch.float32, logits.device)
        stats = _empty(2, torch.float64, logits.device)
        bad = _empty(1, torch.int32, logits.device)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_ce_fwd(L.ptr(logits), L.ptr(target), n, C, int(ignore_index), L.ptr(graw), L.ptr(stats),
                               L.ptr(bad), L.stream()), "spc_ce_fwd")
        if e0 is not None:
            _profiler.end("cross_entropy", e0, 0, (8.0 * C + 8.0) * n, f"C{C} N{n}")
        ctx.save_for_backward(graw, stats)
        _note_bad_flag(bad)
        return (stats[0] / stats[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        graw, stats = ctx.saved_tensors
        n, C = graw.shape
        gout = gout.to(torch.float32).contiguous().view(1)
        dlogits = torch.empty_like(graw)
        L.check(lib.spc_ce_bwd(L.ptr(graw), L.ptr(stats), L.ptr(gout), n, C, L.ptr(dlogits), L.stream()),
                "spc_ce_bwd")
        return dlogits, None, None


class SegHeadFn(torch.autograd.Function):
    """`SegLoss(out.slice(x).F, labels)` (+ `IoUMeter.update`) as one pass over the points (spc_seg_head_fwd):
    inverse-map gather of the voxel logits, class-weighted softmax cross-entropy with ignore index, gradient
    accumulated on the voxel rows, optional per-class counts.  `inverse=None`: rows are the points themselves."""

    @staticmethod
    def forward(ctx, logits, inverse, target, ignore_index, weight, counts):
        lib = L.load()
        logits = _feat(logits)
        if target.dtype != torch.int64:
            raise RuntimeError(f"targets must be int64, got {target.dtype}")
        target = target.contiguous()
        m, C = logits.shape
        n = target.shape[0]
        if target.dim() != 1 or (inverse is None and n != m):
            raise RuntimeError(f"target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}")
        if inverse is not None:
            if inverse.dtype != torch.int32 or inverse.shape != (n,):
                raise RuntimeError("inverse map must be int32 [n]")
            inverse = inverse.contiguous()
        if weight is not None:
            if weight.dtype != torch.float32 or weight.shape != (C,):
                raise RuntimeError(f"class weights must be float32 [{C}]")
            weight = weight.contiguous()
        if counts is not None and (counts.dtype != torch.int64 or counts.shape != (3, C) or not counts.is_contiguous()):
            raise RuntimeError(f"counts must be a contiguous int64 [3, {C}] tensor")
        graw = _empty((m, C), torch.float32, logits.device)
        stats = _empty(2, torch.float64, logits.device)
        bad = _empty(1, torch.int32, logits.device)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_seg_head_fwd(L.ptr(logits), m, L.ptr(inverse), L.ptr(target), n, C, int(ignore_index),
                                     L.ptr(weight), L.ptr(graw), L.ptr(stats), L.ptr(counts), L.ptr(bad), L.stream()),
                "spc_seg_head_fwd")
        if e0 is not None:
            _profiler.end("seg_head", e0, 0, 8.0 * C * n + 12.0 * n, f"C{C} N{n}")
        ctx.save_for_backward(graw, stats)
        _note_bad_flag(bad)
        return (stats[0] / stats[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        graw, stats = ctx.saved_tensors
        m, C = graw.shape
        gout = gout.to(torch.float32).contiguous().view(1)
        dlogits = torch.empty_like(graw)
        L.check(lib.spc_ce_bwd(L.ptr(graw), L.ptr(stats), L.ptr(gout), m, C, L.ptr(dlogits), L.stream()),
                "spc_ce_bwd")
        return dlogits, None, None, None, None, None


# `bad_target` flags of the recent loss kernels (device int32 each).  The kernels zero the gradient of a row whose
# label is outside [0, C) and not ignore_index and raise the flag; reading it needs a host sync, so the training loop
# checks on log steps / at validation end (`raise_on_bad_targets`), where it synchronises anyway.
_bad_flags: list = []


def _note_bad_flag(flag: torch.Tensor) -> None:
    _bad_flags.append(flag)
    if len(_bad_flags) > 4096:  # nobody is checking: keep the newest
        del _bad_flags[:2048]


def raise_on_bad_targets() -> None:
    """Raise if any loss kernel since the last check saw a label outside [0, C) other than ignore_index (what
    F.cross_entropy asserts on).  One host synchronisation."""
    if not _bad_flags:
        return
    flags, _bad_flags[:] = list(_bad_flags), []
    worst = int(torch.stack([f.view(-1)[0] for f in flags]).max().item())
    if worst:
        raise RuntimeError("cross-entropy target outside [0, C) that is not ignore_index"
                           + (" (or an inverse-map entry outside the voxel rows)" if worst == 2 else ""))


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100,
                  weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.cross_entropy(logits, target, weight=..., ignore_index=...) (mean reduction) on the fused kernels.  Targets
    outside [0, C) that are not ignore_index contribute nothing and set the kernel's `bad_target` flag
    (`raise_on_bad_targets`; no host sync here).  More than 64 classes: torch's kernel (the fused ones keep a row in
    registers)."""
    if logits.shape[1] > 64:
        return torch.nn.functional.cross_entropy(logits, target, weight=weight, ignore_index=ignore_index)
    if weight is None:
        return CrossEntropyFn.apply(logits, target, ignore_index)
    return SegHeadFn.apply(logits, None, target, ignore_index, weight, None)


def seg_head(logits: torch.Tensor, inverse: Optional[torch.Tensor], target: torch.Tensor, ignore_index: int = -100,
             weight: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Loss of the voxel logits at the points of the inverse map; `counts` [3, C] int64 accumulates the IoU counts."""
    return SegHeadFn.apply(logits, inverse, target, ignore_index, weight, counts)


class InstanceNormFn(torch.autograd.Function):
    """MinkowskiInstanceNorm: per (batch index, channel) normalisation + [1, C] affine (spc_inst_norm_fwd/bwd)."""

    @staticmethod
    def forward(ctx, x, coords, n_batch, gamma, beta, eps):
        lib = L.load()
        x = _feat(x)
        m, C = x.shape
        dev = x.device
        y = _empty((m, C), torch.float32, dev)
        mean = _empty((n_batch, C), torch.float32, dev)
        rstd = _empty((n_batch, C), torch.float32, dev)
        cnt = _empty(n_batch, torch.int32, dev)
        ws = _empty((n_batch, 2, C), torch.float64, dev)
        g = None if gamma is None else gamma.detach().reshape(-1).contiguous()
        b = None if beta is None else beta.detach().reshape(-1).contiguous()
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_inst_norm_fwd(L.ptr(x), L.ptr(coords), m, C, n_batch, L.ptr(g), L.ptr(b), float(eps), L.ptr(y),
                                      L.ptr(mean), L.ptr(rstd), L.ptr(cnt), L.ptr(ws), L.stream()), "spc_inst_norm_fwd")
        if e0 is not None:
            _profiler.end("inst_norm_fwd", e0, 0, 12.0 * m * C, f"C{C} M{m}")
        ctx.save_for_backward(x, coords, g, mean, rstd, cnt)
        ctx.cfg = (n_batch, gamma is not None and ctx.needs_input_grad[3], beta is not None and ctx.needs_input_grad[4],
                   None if gamma is None else gamma.shape, None if beta is None else beta.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, coords, g, mean, rstd, cnt = ctx.saved_tensors
        n_batch, want_g, want_b, gshape, bshape = ctx.cfg
        dy = _feat(dy)
        m, C = x.shape
        dx = _empty((m, C), torch.float32, x.device)
        sums = _empty((n_batch, 2, C), torch.float64, x.device)
        e0 = _profiler.begin() if _profiler else None
        L.check(lib.spc_inst_norm_bwd(L.ptr(x), L.ptr(dy), L.ptr(coords), m, C, n_batch, L.ptr(g), L.ptr(mean),
                                      L.ptr(rstd), L.ptr(cnt), L.ptr(dx), L.ptr(sums), L.stream()), "spc_inst_norm_bwd")
        if e0 is not None:
            _profiler.end("inst_norm_bwd", e0, 0, 20.0 * m * C, f"C{C} M{m}")
        dgamma = sums[:, 1].sum(0).to(torch.float32).view(gshape) if want_g else None
        dbeta = sums[:, 0].sum(0).to(torch.float32).view(bshape) if want_b else None
        return dx, None, None, dgamma, dbeta, None


# ---------------------------------------------------------------------------
# trilinear interpolation / splat (ME.MinkowskiInterpolation, SparseTensor.interpolate, TensorField.splat)
# ---------------------------------------------------------------------------
def corner_offsets(ts: Sequence[int]):
    """The 8 corners of a lattice cell, k = bx + 2 by + 4 bz (first spatial axis fastest)."""
    return [((k & 1) * int(ts[0]), ((k >> 1) & 1) * int(ts[1]), ((k >> 2) & 1) * int(ts[2])) for k in range(8)]


def interp_corners(query: torch.Tensor, ts: Sequence[int]):
    """query [N,4] float32 (b,x,y,z) -> (lower corner int32 [N,4], weights float32 [8,N])."""
    lib = L.load()
    if query.dtype != torch.float32 or query.dim() != 2 or query.shape[1] != 4:
        raise RuntimeError(f"query coordinates must be float32 [N,4], got {query.dtype} {tuple(query.shape)}")
    query = query.contiguous()
    n = query.shape[0]
    lower = _empty((n, 4), torch.int32, query.device)
    w = _empty((8, n), torch.float32, query.device)
    ts_arr = _I3(*[int(t) for t in ts])
    L.check(lib.spc_interp_corners(L.ptr(query), n, ctypes.cast(ts_arr, ctypes.c_void_p), L.ptr(lower), L.ptr(w),
                                   L.stream()), "spc_interp_corners")
    return lower, w


def interp_map(in_map: CoordMap, query: torch.Tensor):
    """Rows and weights of the 8 voxels of `in_map` around every query point: (idx int32 [8,N], w float32 [8,N])."""
    lower, w = interp_corners(query, in_map.tensor_stride)
    km = build_kernel_map(in_map, CoordMap(lower, None, 0, lower.shape[0], in_map.tensor_stride),
                          corner_offsets(in_map.tensor_stride))
    return km.nbr, w


def _interp_gather(feats, idx, w):
    lib = L.load()
    K, n = idx.shape
    C = feats.shape[1]
    out = _empty((n, C), torch.float32, feats.device)
    e0 = _profiler.begin() if _profiler else None
    L.check(lib.spc_interp_fwd(L.ptr(feats), L.ptr(idx), L.ptr(w), n, C, K, L.ptr(out), L.stream()), "spc_interp_fwd")
    if e0 is not None:
        _profiler.end("interp_gather", e0, 0, 4.0 * C * n * (K + 1) + 8.0 * K * n, f"C{C} N{n}")
    return out


def _interp_scatter(src, idx, w, m):
    lib = L.load()
    K, n = idx.shape
    C = src.shape[1]
    out = _empty((m, C), torch.float32, src.device)
    e0 = _profiler.begin() if _profiler else None
    L.check(lib.spc_interp_bwd(L.ptr(src), L.ptr(idx), L.ptr(w), n, m, C, K, L.ptr(out), L.stream()), "spc_interp_bwd")
    if e0 is not None:
        _profiler.end("interp_scatter", e0, 0, 4.0 * C * n * (K + 1) + 8.0 * K * n, f"C{C} N{n}")
    return out


class InterpolateFn(torch.autograd.Function):
    """out[j] = sum_k w[k,j] * feats[idx[k,j]] — voxel features at the query points."""

    @staticmethod
    def forward(ctx, feats, idx, w):
        feats = _feat(feats)
        ctx.save_for_backward(idx, w)
        ctx.m = feats.shape[0]
        return _interp_gather(feats, idx, w)

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.saved_tensors
        return _interp_scatter(_feat(g), idx, w, ctx.m), None, None


class SplatFn(torch.autograd.Function):
    """out[idx[k,j]] += w[k,j] * feats[j] — point features spread over their 8 voxels (transpose of InterpolateFn)."""

    @staticmethod
    def forward(ctx, feats, idx, w, m):
        feats = _feat(feats)
        ctx.save_for_backward(idx, w)
        return _interp_scatter(feats, idx, w, m)

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.saved_tensors
        return _interp_gather(_feat(g), idx, w), None, None, None


def sgd_step(param, grad, buf, lr, momentum, weight_decay, grad_scale, first_step):
    lib = L.load()
    invalidate_packed_weights()  # the kernel rewrites every parameter through a raw pointer
    L.check(lib.spc_sgd_step(L.ptr(param), L.ptr(grad), L.ptr(buf), param.numel(), float(lr), float(momentum),
                             float(weight_decay), float(grad_scale), int(first_step), L.stream()), "spc_sgd_step")
    if batch_repack and param.is_cuda:
        repack_all()   # every layer's packed weight images, one launch
