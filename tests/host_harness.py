"""HOST-LOGIC TEST HARNESS — not a backend, not a fallback, never imported by the product.

The Python layer above the C ABI (coordinate manager, key rules, kernel-map caching, lazy BatchNorm fusion, channel
padding, bf16 side-copy bookkeeping, autograd wiring, the training loop) is ordinary host code, but it can only run
when `libsparseconv_b200.so` answers its calls — on a GPU.  GPU time is scarce, so this module lets `-m "not gpu"`
tests drive that host code on CPU tensors: `install(monkeypatch)` swaps `lib.load()` for `FakeLib`, an object with
the entry points of include/sparseconv_b200.h whose bodies are the ORACLE's numpy restatements applied to the raw
pointers the host code passes.  What such a test proves is that the host code passes the right pointers, sizes, maps
and flags in the right order; it proves nothing about the CUDA kernels (tests/test_gpu_*.py do, through the real
library).  The product keeps refusing CPU tensors (`lib.ptr`, `TensorField`): the patches below live only inside a
pytest `monkeypatch` scope.

Entry points not emulated raise `NotImplementedError` naming themselves, so a test that strays onto an un-emulated
path fails loudly instead of reading garbage.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from oracle import ref_ops as R

SRC_FLOAT, SRC_INT, SRC_STRIDE = 0, 1, 2
PREC_BF16 = 2


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def view(p, shape, dtype):
    """numpy view of host memory at raw address `p` (None for NULL)."""
    a = _addr(p)
    if a == 0:
        return None
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(a)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def rows_view(p, m, C, pitch, dtype=np.float32):
    """[m, C] view of rows stored `pitch` elements apart (reads only the valid elements)."""
    a = _addr(p)
    if m == 0:
        return np.zeros((0, C), dtype)
    item = np.dtype(dtype).itemsize
    span = ((m - 1) * pitch + C) * item
    buf = (ctypes.c_char * span).from_address(a)
    base = np.frombuffer(buf, dtype=dtype, count=(m - 1) * pitch + C)
    return np.lib.stride_tricks.as_strided(base, shape=(m, C), strides=(pitch * item, item), writeable=False)


def bf16_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    """round to nearest even, as the kernels do"""
    t = torch.from_numpy(np.array(x, np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16)


def feat_rows(p, m, C, precision):
    """fp32 [m, C] contents of a feature operand (bf16 rows when precision == SPC_PREC_BF16)."""
    if precision == PREC_BF16:
        return bf16_to_f32(view(p, (m, C), np.uint16))
    return view(p, (m, C), np.float32)


class FakeLib:
    def __init__(self):
        self.tables = {}        # table address -> int32 coords [M, 4] of the map it indexes
        self.calls = []         # names, in order
        self.launches = 0

    def __getattr__(self, name):
        if name.startswith("spc_"):
            def missing(*a, **k):
                raise NotImplementedError(f"{name} is not emulated by tests/host_harness.py")
            return missing
        raise AttributeError(name)

    def _called(self, name):
        self.calls.append(name)
        self.launches += 1

    # ---- misc -----------------------------------------------------------------------------------------------
    def spc_abi_version(self):
        return 1

    def spc_last_error(self):
        return b"host harness"

    def spc_launch_count(self):
        return self.launches

    def spc_table_slots(self, n):
        s = 16
        while s < 2 * max(int(n), 1):
            s *= 2
        return s

    def spc_coords_insert_workspace(self, n):
        return 256

    def spc_conv_workspace(self, K, c_in, c_out, precision):
        return 256

    def spc_bn_workspace(self, m, C):
        return 256

    def spc_pairs_workspace(self, m_out, K):
        return 256

    # ---- coordinates ------------------------------------------------------------------------------------------
    def spc_coords_insert(self, src, n, kind, ts, slots, n_slots, out_coords, out_first, out_inverse, out_count,
                          status, ws, ws_bytes, stream):
        self._called("spc_coords_insert")
        ts = tuple(int(v) for v in view(ts, 3, np.int32))
        st = view(status, 2, np.int32)
        if n == 0:
            st[:] = 0
            self.tables[_addr(slots)] = np.zeros((0, 4), np.int32)
            return 0
        if kind == SRC_FLOAT:
            q = R.quantize_np(view(src, (n, 4), np.float32), ts)
        elif kind == SRC_INT:
            q = view(src, (n, 4), np.int32).copy()
        else:
            q = view(src, (n, 4), np.int32).copy()
            t = np.array((1,) + ts, np.int64)
            q = (np.floor_divide(q.astype(np.int64), t) * t).astype(np.int32)
        bad = (q[:, 0] < 0) | (q[:, 0] > 1022) | (np.abs(q[:, 1:].astype(np.int64)) > 131071).any(1)
        uc, ui, inv = R.unique_first_np(q)
        m = uc.shape[0]
        view(out_coords, (n, 4), np.int32)[:m] = uc
        view(out_first, n, np.int32)[:m] = ui
        view(out_inverse, n, np.int32)[:] = inv
        view(out_count, n, np.int32)[:m] = np.bincount(inv, minlength=m)
        st[0], st[1] = m, int(bad.any())
        self.tables[_addr(slots)] = uc.copy()
        return 0

    def spc_coords_insert_dev(self, src, n, n_dev, kind, ts, slots, n_slots, out_coords, out_first, out_inverse,
                              out_count, status, ws, ws_bytes, stream):
        """rows = min(n, *n_dev): the row count of the parent level is read from ITS status word (stride pyramid)."""
        rows = n if not _addr(n_dev) else min(int(n), int(view(n_dev, 2, np.int32)[0]))
        rc = self.spc_coords_insert(src, rows, kind, ts, slots, n_slots, out_coords, out_first, out_inverse, out_count,
                                    status, ws, ws_bytes, stream)
        self.calls[-1] = "spc_coords_insert_dev"
        return rc

    def spc_kernel_map(self, in_slots, in_n_slots, out_coords, m_out, offsets, K, nbr, tap_count, stream):
        self._called("spc_kernel_map")
        in_coords = self.tables[_addr(in_slots)]
        offs = [tuple(int(v) for v in o) for o in view(offsets, (K, 3), np.int32)]
        out = view(nbr, (K, m_out), np.int32)
        tc = view(tap_count, K, np.int32)
        if m_out == 0:
            tc[:] = 0
            return 0
        res = R.kernel_map_np(in_coords, view(out_coords, (m_out, 4), np.int32), offs)
        out[:] = res
        tc[:] = (res >= 0).sum(1)
        return 0

    def spc_kernel_map_sym(self, slots, n_slots, coords, m, offsets, K, nbr, tap_count, stream):
        offs = view(offsets, (K, 3), np.int32)
        assert K % 2 == 1 and (offs == -offs[::-1]).all(), "offsets are not centrally symmetric"
        return self.spc_kernel_map(slots, n_slots, coords, m, offsets, K, nbr, tap_count, stream)

    def spc_kernel_map_transpose(self, nbr, m_out, m_in, K, nbr_t, stream):
        self._called("spc_kernel_map_transpose")
        view(nbr_t, (K, m_in), np.int32)[:] = R.transpose_dense(view(nbr, (K, m_out), np.int32), m_in)
        return 0

    def spc_row_masks(self, nbr, m, K, row_mask, executed_dev, stream):
        self._called("spc_row_masks")
        nb = view(nbr, (K, m), np.int32)
        bits = np.zeros(m, np.uint32)
        for k in range(K):
            bits |= (nb[k] >= 0).astype(np.uint32) << np.uint32(k)
        view(row_mask, m, np.uint32)[:] = bits
        pad = np.concatenate([bits, np.zeros((-m) % 128, np.uint32)]).reshape(-1, 128)
        tile = np.bitwise_or.reduce(pad, axis=1)
        view(executed_dev, 1, np.int64)[0] = int(sum(bin(int(t)).count("1") for t in tile))
        return 0

    def spc_table_relabel(self, slots, n_slots, pos, stream):
        """the fake table IS the coordinate array in row order: old row r moves to pos[r]"""
        self._called("spc_table_relabel")
        coords = self.tables[_addr(slots)]
        p = view(pos, coords.shape[0], np.int32)
        new = np.empty_like(coords)
        new[p] = coords
        self.tables[_addr(slots)] = new
        return 0

    def spc_tile_mask(self, nbr, m, K, mask, stream):
        self._called("spc_tile_mask")
        n_tiles = (m + 127) // 128
        a = view(nbr, (K, m), np.int32)
        out = view(mask, max(n_tiles, 1), np.uint32)
        for t in range(n_tiles):
            bits = 0
            for k in range(K):
                if (a[k, t * 128:(t + 1) * 128] >= 0).any():
                    bits |= 1 << k
            out[t] = bits
        return 0

    def spc_kernel_map_pairs(self, nbr, m_out, K, cap, pairs, tap_off, ws, ws_bytes, stream):
        self._called("spc_kernel_map_pairs")
        a = view(nbr, (K, m_out), np.int32)
        p = view(pairs, (2, cap), np.int32)
        off = view(tap_off, K + 1, np.int32)
        pos = 0
        for k in range(K):
            off[k] = pos
            o = np.nonzero(a[k] >= 0)[0]
            p[0, pos:pos + o.size] = a[k, o]
            p[1, pos:pos + o.size] = o
            pos += o.size
        off[K] = pos
        return 0

    # ---- feature rows -----------------------------------------------------------------------------------------
    def spc_segment_reduce(self, feats, inverse, count, n, m, C, mode, out, stream):
        self._called("spc_segment_reduce")
        o = view(out, (m, C), np.float32)
        o[:] = 0
        inv = view(inverse, n, np.int32)
        f = view(feats, (n, C), np.float32)
        if mode == 2:                                               # random subsample == first point of the voxel
            first = np.full(m, n, np.int64)
            np.minimum.at(first, inv, np.arange(n))
            o[:] = f[first]
            return 0
        acc = np.zeros((m, C), np.float64)
        np.add.at(acc, inv, f.astype(np.float64))
        if mode == 0:
            acc /= np.maximum(view(count, m, np.int32), 1)[:, None]
        o[:] = acc
        return 0

    def spc_gather_rows(self, src, index, count, n, C, out, stream):
        self._called("spc_gather_rows")
        idx = view(index, n, np.int32)
        m = int(idx.max()) + 1 if n else 0
        s = view(src, (m, C), np.float32)
        o = view(out, (n, C), np.float32)
        if n:
            o[:] = s[idx]
            if _addr(count):
                o[:] = o / view(count, m, np.int32)[idx][:, None]
        return 0

    def spc_scatter_add_rows(self, src, index, n, m, C, out, stream):
        self._called("spc_scatter_add_rows")
        acc = np.zeros((m, C), np.float64)
        np.add.at(acc, view(index, n, np.int32), view(src, (n, C), np.float32).astype(np.float64))
        view(out, (m, C), np.float32)[:] = acc
        return 0

    # ---- convolution ------------------------------------------------------------------------------------------
    def spc_to_bf16(self, src, rows, c_src, src_pitch, c_dst, dst, stream):
        self._called("spc_to_bf16")
        d = view(dst, (rows, c_dst), np.uint16)
        d[:] = 0
        d[:, :c_src] = f32_to_bf16(rows_view(src, rows, c_src, src_pitch))
        return 0

    @staticmethod
    def _apply_tile_mask(a, mask, m, K):
        """The tensor-core kernels skip offset k for a 128-row tile whose mask bit k is clear (spc_tile_mask, or a
        mask restricted to the offsets of a weight-sparse convolution)."""
        if not _addr(mask) or K > 32 or m == 0:
            return a
        tm = view(mask, ((m + 127) // 128,), np.int32).astype(np.int64) & 0xFFFFFFFF
        rows_tile = np.arange(m) // 128
        a = a.copy()
        for k in range(K):
            a[k, ((tm[rows_tile] >> k) & 1) == 0] = -1
        return a

    def spc_conv_fwd(self, x, w, bias, nbr, mask, m_in, m_out, c_in, c_out, K, precision, out, ws, ws_bytes, stream):
        self._called("spc_conv_fwd")
        xin = feat_rows(x, m_in, c_in, precision).astype(np.float64)
        W = view(w, (K, c_in, c_out), np.float32).astype(np.float64)
        a = view(nbr, (K, m_out), np.int32)
        a = self._apply_tile_mask(a, mask, m_out, K)
        acc = np.zeros((m_out, c_out), np.float64)
        for k in range(K):
            o = np.nonzero(a[k] >= 0)[0]
            if o.size:
                acc[o] += xin[a[k, o]] @ W[k]
        if _addr(bias):
            acc += view(bias, c_out, np.float32)
        view(out, (m_out, c_out), np.float32)[:] = acc
        return 0

    def spc_conv_dgrad(self, dout, w, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, ws, ws_bytes, stream):
        self._called("spc_conv_dgrad")
        g = feat_rows(dout, m_out, c_out, precision).astype(np.float64)
        W = view(w, (K, c_in, c_out), np.float32).astype(np.float64)
        a = view(nbr_t, (K, m_in), np.int32)
        a = self._apply_tile_mask(a, mask_t, m_in, K)
        acc = np.zeros((m_in, c_in), np.float64)
        for k in range(K):
            i = np.nonzero(a[k] >= 0)[0]
            if i.size:
                acc[i] += g[a[k, i]] @ W[k].T
        view(din, (m_in, c_in), np.float32)[:] = acc
        return 0

    def spc_conv_wgrad(self, x, dout, nbr, mask, m_in, m_out, c_in, c_out, K, precision, dw, ws, ws_bytes, stream):
        self._called("spc_conv_wgrad")
        xin = feat_rows(x, m_in, c_in, precision).astype(np.float64)
        g = feat_rows(dout, m_out, c_out, precision).astype(np.float64)
        a = view(nbr, (K, m_out), np.int32)
        d = view(dw, (K, c_in, c_out), np.float32)
        for k in range(K):
            o = np.nonzero(a[k] >= 0)[0]
            d[k] = xin[a[k, o]].T @ g[o] if o.size else 0.0
        return 0

    def spc_conv_wgrad_acc(self, x, dout, nbr, mask, m_in, m_out, c_in, c_out, K, precision, dw, accumulate, stream):
        if not accumulate:
            return self.spc_conv_wgrad(x, dout, nbr, mask, m_in, m_out, c_in, c_out, K, precision, dw, None, 0, stream)
        self._called("spc_conv_wgrad_acc")
        assert self.spc_conv_tensor_core(2, K, c_in, c_out, precision), "accumulate needs a tensor-core shape"
        xin = feat_rows(x, m_in, c_in, precision).astype(np.float64)
        g = feat_rows(dout, m_out, c_out, precision).astype(np.float64)
        a = view(nbr, (K, m_out), np.int32)
        d = view(dw, (K, c_in, c_out), np.float32)
        for k in range(K):
            o = np.nonzero(a[k] >= 0)[0]
            if o.size:
                d[k] += (xin[a[k, o]].T @ g[o]).astype(np.float32)
        return 0

    # the tensor-core routing rules of conv_api.cu (what: 0 fwd, 1 dgrad, 2 wgrad)
    def spc_conv_tensor_core(self, what, K, c_in, c_out, precision):
        if precision == 0 or K > 32:
            return 0
        fwd_ok = lambda ck, cn: ck >= 32 and ck % 32 == 0 and cn % 16 == 0   # noqa: E731
        if what == 0:
            return int(fwd_ok(c_in, c_out))
        if what == 1:
            return int(fwd_ok(c_out, c_in))
        return int(K * c_in <= 128 * 128 and c_in >= 32 and c_in % 32 == 0 and 32 <= c_out <= 1024 and c_out % 32 == 0)

    def spc_conv_packed_bytes(self, K, c_in, c_out):
        return (K * c_in * c_out * 4 + 1023) // 1024 * 1024

    def spc_conv_path_counts(self, out3, reset):
        return None

    def spc_conv_pack_weights(self, w, K, c_in, c_out, dgrad, precision, packed, stream):
        """The harness's "packed image" is the fp32 kernel itself (the real one is a swizzled slab layout)."""
        self._called("spc_conv_pack_weights")
        assert _addr(packed) % 1024 == 0
        src = view(w, (K, c_in, c_out), np.float32)
        view(packed, (K, c_in, c_out), np.float32)[:] = src[::-1] if dgrad == 2 else src   # 2: offsets reversed
        return 0

    def spc_conv_pack_weights_batch(self, desc, n_layers, stream):
        self._called("spc_conv_pack_weights_batch")
        d = view(desc, (n_layers, 8), np.int64)
        for w, packed, K, ck, cn, flags, bf16, _ in d.tolist():
            c_in, c_out = (cn, ck) if flags & 1 else (ck, cn)
            src = view(w, (K, c_in, c_out), np.float32)
            view(packed, (K, c_in, c_out), np.float32)[:] = src[::-1] if flags & 2 else src
        return 0

    def spc_conv_fwd_packed(self, x, wp, bias, nbr, mask, m_in, m_out, c_in, c_out, K, precision, out, stream):
        assert self.spc_conv_tensor_core(0, K, c_in, c_out, precision)
        return self.spc_conv_fwd(x, wp, bias, nbr, mask, m_in, m_out, c_in, c_out, K, precision, out, None, 0, stream)

    def spc_conv_fwd_packed_stats(self, x, wp, bias, nbr, mask, m_in, m_out, c_in, c_out, K, precision, out, bn_sums,
                                  stats_fused, stream):
        rc = self.spc_conv_fwd_packed(x, wp, bias, nbr, mask, m_in, m_out, c_in, c_out, K, precision, out, stream)
        fused = int(_addr(bn_sums) != 0 and not _addr(bias) and c_out <= 128 and m_out >= 4096)  # (large maps only)
        if fused:
            o = view(out, (m_out, c_out), np.float32).astype(np.float64)
            s = view(bn_sums, 2 * c_out, np.float64)
            s[:c_out] = o.sum(0)
            s[c_out:] = (o * o).sum(0)
        if stats_fused is not None:
            stats_fused._obj.value = fused
        return rc

    def spc_bn_finalize(self, sums, m, C, mean, var, run_mean, run_var, momentum, tracked, stream):
        self._called("spc_bn_finalize")
        s = view(sums, 2 * C, np.float64)
        mu = s[:C] / m
        v = np.maximum(s[C:] / m - mu * mu, 0.0)
        view(mean, C, np.float32)[:] = mu
        view(var, C, np.float32)[:] = v
        if _addr(run_mean):
            rm, rv = view(run_mean, C, np.float32), view(run_var, C, np.float32)
            rm[:] = (1 - momentum) * rm + momentum * mu
            rv[:] = (1 - momentum) * rv + momentum * (v * m / max(m - 1, 1))
        if _addr(tracked):
            view(tracked, 1, np.int64)[:] += 1
        return 0

    def spc_conv_dgrad_packed(self, dout, wp, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, stream):
        assert self.spc_conv_tensor_core(1, K, c_in, c_out, precision)
        return self.spc_conv_dgrad(dout, wp, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, None, 0, stream)

    def spc_conv_dgrad_packed_acc(self, dout, wp, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, accumulate,
                                  stream):
        if not accumulate:
            return self.spc_conv_dgrad_packed(dout, wp, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, stream)
        before = view(din, (m_in, c_in), np.float32).copy()
        rc = self.spc_conv_dgrad_packed(dout, wp, nbr_t, mask_t, m_in, m_out, c_in, c_out, K, precision, din, stream)
        view(din, (m_in, c_in), np.float32)[:] += before
        return rc

    # ---- batch norm / elementwise -----------------------------------------------------------------------------
    def spc_bn_stats_tracked(self, x, m, C, mean, var, run_mean, run_var, momentum, tracked, ws, ws_bytes, stream):
        rc = self.spc_bn_stats(x, m, C, mean, var, run_mean, run_var, momentum, ws, ws_bytes, stream)
        if _addr(tracked):
            view(tracked, 1, np.int64)[:] += 1
        return rc

    def spc_bn_bwd_acc(self, x, y, y_bf16, dy, dy_pitch, mean, var, gamma, beta, m, C, eps, relu, training, dx, dx_bf16,
                       dres, dgamma, dbeta, accumulate, ws, ws_bytes, stream):
        if not accumulate:
            return self.spc_bn_bwd(x, y, y_bf16, dy, dy_pitch, mean, var, gamma, m, C, eps, relu, training, dx, dx_bf16,
                                   dres, dgamma, dbeta, ws, ws_bytes, stream, beta=beta)
        g0, b0 = view(dgamma, C, np.float32).copy(), view(dbeta, C, np.float32).copy()
        rc = self.spc_bn_bwd(x, y, y_bf16, dy, dy_pitch, mean, var, gamma, m, C, eps, relu, training, dx, dx_bf16,
                             dres, dgamma, dbeta, ws, ws_bytes, stream, beta=beta)
        view(dgamma, C, np.float32)[:] += g0
        view(dbeta, C, np.float32)[:] += b0
        return rc

    def spc_bn_stats(self, x, m, C, mean, var, run_mean, run_var, momentum, ws, ws_bytes, stream):
        self._called("spc_bn_stats")
        xv = view(x, (m, C), np.float32).astype(np.float64)
        mu, v = xv.mean(0), xv.var(0)
        view(mean, C, np.float32)[:] = mu
        view(var, C, np.float32)[:] = v
        if _addr(run_mean):
            rm, rv = view(run_mean, C, np.float32), view(run_var, C, np.float32)
            rm[:] = (1 - momentum) * rm + momentum * mu
            rv[:] = (1 - momentum) * rv + momentum * (v * m / max(m - 1, 1))
        return 0

    def spc_bn_apply(self, x, mean, var, gamma, beta, res, m, C, eps, relu, y, y_bf16, stream):
        self._called("spc_bn_apply")
        xv = view(x, (m, C), np.float32).astype(np.float64)
        o = (xv - view(mean, C, np.float32)) / np.sqrt(view(var, C, np.float32).astype(np.float64) + eps)
        if _addr(gamma):
            o = o * view(gamma, C, np.float32)
        if _addr(beta):
            o = o + view(beta, C, np.float32)
        if _addr(res):
            o = o + view(res, (m, C), np.float32)
        if relu:
            o = np.maximum(o, 0)
        if _addr(y):
            view(y, (m, C), np.float32)[:] = o
        if _addr(y_bf16):
            view(y_bf16, (m, C), np.uint16)[:] = f32_to_bf16(o.astype(np.float32))
        return 0

    def spc_bn_bwd(self, x, y, y_bf16, dy, dy_pitch, mean, var, gamma, m, C, eps, relu, training, dx, dx_bf16, dres,
                   dgamma, dbeta, ws, ws_bytes, stream, beta=None):
        self._called("spc_bn_bwd")
        xv = view(x, (m, C), np.float32).astype(np.float64)
        g = np.array(rows_view(dy, m, C, dy_pitch), np.float64)
        if relu == 1:
            yv = view(y, (m, C), np.float32) if _addr(y) else bf16_to_f32(view(y_bf16, (m, C), np.uint16))
            g = g * (yv > 0)
        elif relu == 2:   # mask recomputed from x with the forward affine (no residual before the ReLU)
            x32 = view(x, (m, C), np.float32)
            ga = view(gamma, C, np.float32) if _addr(gamma) else np.ones(C, np.float32)
            be = view(beta, C, np.float32) if _addr(beta) else np.zeros(C, np.float32)
            sc = (ga / np.sqrt(view(var, C, np.float32) + np.float32(eps))).astype(np.float32)
            sh = (be - view(mean, C, np.float32) * sc).astype(np.float32)
            g = g * ((x32 * sc + sh) > 0)
        if _addr(dres):
            view(dres, (m, C), np.float32)[:] = g
        rstd = 1.0 / np.sqrt(view(var, C, np.float32).astype(np.float64) + eps)
        xh = (xv - view(mean, C, np.float32)) * rstd
        view(dbeta, C, np.float32)[:] = g.sum(0)
        view(dgamma, C, np.float32)[:] = (g * xh).sum(0)
        sc = rstd * (view(gamma, C, np.float32) if _addr(gamma) else 1.0)
        d = sc * (g - g.mean(0) - xh * (g * xh).mean(0)) if training else sc * g
        if _addr(dx):
            view(dx, (m, C), np.float32)[:] = d
        if _addr(dx_bf16):
            view(dx_bf16, (m, C), np.uint16)[:] = f32_to_bf16(d.astype(np.float32))
        return 0

    def spc_copy_rows(self, src, src_pitch, dst, dst_pitch, row_bytes, rows, stream):
        self._called("spc_copy_rows")
        for r in range(int(rows)):
            ctypes.memmove(_addr(dst) + r * int(dst_pitch), _addr(src) + r * int(src_pitch), int(row_bytes))
        return 0

    def spc_relu_fwd(self, x, n, y, stream):
        self._called("spc_relu_fwd")
        view(y, n, np.float32)[:] = np.maximum(view(x, n, np.float32), 0)
        return 0

    def spc_relu_bwd(self, y, dy, n, dx, stream):
        self._called("spc_relu_bwd")
        view(dx, n, np.float32)[:] = view(dy, n, np.float32) * (view(y, n, np.float32) > 0)
        return 0

    def spc_add(self, a, b, n, y, stream):
        self._called("spc_add")
        view(y, n, np.float32)[:] = view(a, n, np.float32) + view(b, n, np.float32)
        return 0

    # ---- pooling ----------------------------------------------------------------------------------------------
    def spc_pool_fwd(self, x, nbr, m_out, C, K, avg, out, stream):
        self._called("spc_pool_fwd")
        a = view(nbr, (K, m_out), np.int32)
        m_in = int(a.max()) + 1 if a.size else 0
        xin = view(x, (m_in, C), np.float32).astype(np.float64)
        acc = np.zeros((m_out, C), np.float64)
        cnt = np.zeros(m_out)
        for k in range(K):
            o = np.nonzero(a[k] >= 0)[0]
            acc[o] += xin[a[k, o]]
            cnt[o] += 1
        if avg:
            acc /= np.maximum(cnt, 1)[:, None]
        view(out, (m_out, C), np.float32)[:] = acc
        return 0

    def spc_pool_max_fwd(self, x, nbr, m_out, C, K, out, arg, stream):
        self._called("spc_pool_max_fwd")
        a = view(nbr, (K, m_out), np.int32)
        m_in = int(a.max()) + 1 if a.size else 0
        o, g = R.pool_max_np(view(x, (m_in, C), np.float32), a)
        view(out, (m_out, C), np.float32)[:] = o
        view(arg, (m_out, C), np.int32)[:] = g
        return 0

    def spc_pool_max_bwd(self, dout, arg, m_out, m_in, C, din, stream):
        self._called("spc_pool_max_bwd")
        d = view(din, (m_in, C), np.float32)
        d[:] = 0
        a, g = view(arg, (m_out, C), np.int32), view(dout, (m_out, C), np.float32)
        for c in range(C):
            ok = a[:, c] >= 0
            np.add.at(d[:, c], a[ok, c], g[ok, c])
        return 0

    def spc_global_pool_fwd(self, x, coords, m, C, n_batch, avg, out, cnt, stream):
        self._called("spc_global_pool_fwd")
        b = view(coords, (m, 4), np.int32)[:, 0]
        xin = view(x, (m, C), np.float32).astype(np.float64)
        acc = np.zeros((n_batch, C), np.float64)
        np.add.at(acc, b, xin)
        c = np.bincount(b, minlength=n_batch).astype(np.int32)
        view(cnt, n_batch, np.int32)[:] = c
        if avg:
            acc /= np.maximum(c, 1)[:, None]
        view(out, (n_batch, C), np.float32)[:] = acc
        return 0

    def spc_global_pool_bwd(self, dout, coords, cnt, m, C, n_batch, avg, din, stream):
        self._called("spc_global_pool_bwd")
        b = view(coords, (m, 4), np.int32)[:, 0]
        g = view(dout, (n_batch, C), np.float32)[b].astype(np.float64)
        if avg:
            g /= view(cnt, n_batch, np.int32)[b][:, None]
        view(din, (m, C), np.float32)[:] = g
        return 0

    def spc_global_max_fwd(self, x, coords, m, C, n_batch, out, arg, ws, ws_bytes, stream):
        self._called("spc_global_max_fwd")
        o, a = R.global_max_np(view(x, (m, C), np.float32), view(coords, (m, 4), np.int32)[:, 0], n_batch)
        view(out, (n_batch, C), np.float32)[:] = o
        view(arg, (n_batch, C), np.int32)[:] = a
        return 0

    # ---- loss / metrics / optimiser ---------------------------------------------------------------------------
    def spc_ce_fwd(self, logits, target, n, C, ignore, graw, stats, bad, stream):
        return self.spc_seg_head_fwd(logits, n, None, target, n, C, ignore, None, graw, stats, None, bad, stream,
                                     name="spc_ce_fwd")

    def spc_seg_head_fwd(self, logits, m, inverse, target, n, C, ignore, weight, graw, stats, counts, bad, stream,
                         name="spc_seg_head_fwd"):
        self._called(name)
        lg = torch.from_numpy(view(logits, (m, C), np.float32).astype(np.float64))
        inv = view(inverse, n, np.int32) if _addr(inverse) else None
        tgt = view(target, n, np.int64)
        w = torch.from_numpy(view(weight, C, np.float32).astype(np.float64)) if _addr(weight) else None
        rows = torch.arange(m) if inv is None else torch.from_numpy(inv.astype(np.int64))
        valid = tgt != ignore
        view(bad, 1, np.int32)[0] = int(((tgt[valid] < 0) | (tgt[valid] >= C)).any())
        t = torch.from_numpy(np.where(valid, tgt, 0).astype(np.int64))
        lp = torch.log_softmax(lg[rows], 1)
        wy = (w[t] if w is not None else torch.ones(n, dtype=torch.float64)) * torch.from_numpy(valid)
        st = view(stats, 2, np.float64)
        st[0] = float(-(wy * lp[torch.arange(n), t]).sum())
        st[1] = float(wy.sum())
        g = torch.exp(lp)
        g[torch.arange(n), t] -= 1
        g = g * wy[:, None]
        acc = torch.zeros((m, C), dtype=torch.float64).index_add_(0, rows, g)
        view(graw, (m, C), np.float32)[:] = acc.numpy()
        if _addr(counts):
            c = view(counts, (3, C), np.int64)
            c += R.iou_counts_np(lg[rows].numpy(), tgt, C, ignore)
        return 0

    def spc_ce_bwd(self, graw, stats, gout, n, C, dlogits, stream):
        self._called("spc_ce_bwd")
        st = view(stats, 2, np.float64)
        sc = float(view(gout, 1, np.float32)[0]) / st[1] if st[1] > 0 else 0.0
        view(dlogits, (n, C), np.float32)[:] = view(graw, (n, C), np.float32) * sc
        return 0

    def spc_seg_metrics(self, logits, target, n, C, ignore, counts, stream):
        self._called("spc_seg_metrics")
        c = view(counts, (3, C), np.int64)
        c += R.iou_counts_np(view(logits, (n, C), np.float32), view(target, n, np.int64), C, ignore)
        return 0

    def spc_sgd_step(self, param, grad, buf, n, lr, momentum, wd, grad_scale, first_step, stream):
        self._called("spc_sgd_step")
        p, g, b = view(param, n, np.float32), view(grad, n, np.float32), view(buf, n, np.float32)
        d = g * np.float32(grad_scale) + np.float32(wd) * p
        b[:] = d if first_step else np.float32(momentum) * b + d
        p[:] = p - np.float32(lr) * b
        return 0

    def spc_plenoxel_decode(self, links, is64, n, reso, batch_index, affine12, sh_u8, C, sh_scale, sh_min, out_coords,
                            out_feats, stream):
        self._called("spc_plenoxel_decode")
        r = [int(v) for v in view(reso, 3, np.int32)]
        ln = view(links, n, np.int64 if is64 else np.int32).astype(np.int64)
        aff = list(view(affine12, 12, np.float32)) if _addr(affine12) else None
        sh = view(sh_u8, (n, C), np.uint8) if C else np.zeros((n, 0), np.uint8)
        c, f = R.plenoxel_decode_np(ln, sh, np.float32(sh_scale), np.float32(sh_min), r, batch_index=batch_index, affine=aff)
        view(out_coords, (n, 4), np.float32)[:] = c
        if C:
            view(out_feats, (n, C), np.float32)[:] = f
        return 0

    def spc_plenoxel_decode_rows(self, links, is64, rows, n_rows, reso, batch_index, affine12, sh_u8, C, sh_scale, sh_min,
                                 out_coords, out_feats, stream):
        self._called("spc_plenoxel_decode_rows")
        r = [int(v) for v in view(reso, 3, np.int32)]
        rw = view(rows, n_rows, np.int32).astype(np.int64)
        n_rec = int(rw.max()) + 1 if n_rows else 0
        ln = view(links, n_rec, np.int64 if is64 else np.int32).astype(np.int64)[rw]
        aff = list(view(affine12, 12, np.float32)) if _addr(affine12) else None
        sh = view(sh_u8, (n_rec, C), np.uint8)[rw] if C else np.zeros((n_rows, 0), np.uint8)
        c, f = R.plenoxel_decode_np(ln, sh, np.float32(sh_scale), np.float32(sh_min), r, batch_index=batch_index, affine=aff)
        view(out_coords, (n_rows, 4), np.float32)[:] = c
        if C:
            view(out_feats, (n_rows, C), np.float32)[:] = f
        return 0

    def spc_plenoxel_crop_workspace(self, n):
        return 256

    def spc_plenoxel_crop_select(self, links, is64, rows, n, reso, affine12, u3, size3, out_rows, result2, ws, ws_bytes,
                                 stream):
        self._called("spc_plenoxel_crop_select")
        r = [int(v) for v in view(reso, 3, np.int32)]
        rw = view(rows, n, np.int32).astype(np.int64) if _addr(rows) else np.arange(n, dtype=np.int64)
        n_rec = int(rw.max()) + 1 if n else 0
        ln = view(links, n_rec, np.int64 if is64 else np.int32).astype(np.int64)[rw]
        aff = list(view(affine12, 12, np.float32)) if _addr(affine12) else None
        c, _ = R.plenoxel_decode_np(ln, np.zeros((n, 0), np.uint8), 1.0, 0.0, r, affine=aff)
        keep, fits = R.random_crop_select_np(c[:, 1:], view(u3, 3, np.float32), view(size3, 3, np.float32))
        view(out_rows, n, np.int32)[:len(keep)] = rw[keep]
        view(result2, 2, np.int32)[:] = (len(keep), int(fits))
        return 0

    # ---- instance norm / interpolation ------------------------------------------------------------------------
    def spc_inst_norm_fwd(self, x, coords, m, C, n_batch, gamma, beta, eps, y, mean, rstd, cnt, ws, stream):
        self._called("spc_inst_norm_fwd")
        b = view(coords, (m, 4), np.int32)[:, 0]
        xv = view(x, (m, C), np.float32).astype(np.float64)
        mu, rs = np.zeros((n_batch, C)), np.zeros((n_batch, C))
        c = np.bincount(b, minlength=n_batch)
        for i in range(n_batch):
            if c[i]:
                mu[i] = xv[b == i].mean(0)
                rs[i] = 1.0 / np.sqrt(xv[b == i].var(0) + eps)
        o = (xv - mu[b]) * rs[b]
        if _addr(gamma):
            o = o * view(gamma, C, np.float32)
        if _addr(beta):
            o = o + view(beta, C, np.float32)
        view(y, (m, C), np.float32)[:] = o
        view(mean, (n_batch, C), np.float32)[:] = mu
        view(rstd, (n_batch, C), np.float32)[:] = rs
        view(cnt, n_batch, np.int32)[:] = c
        return 0

    def spc_inst_norm_bwd(self, x, dy, coords, m, C, n_batch, gamma, mean, rstd, cnt, dx, sums, stream):
        self._called("spc_inst_norm_bwd")
        b = view(coords, (m, 4), np.int32)[:, 0]
        xv, g = view(x, (m, C), np.float32).astype(np.float64), view(dy, (m, C), np.float32).astype(np.float64)
        mu, rs = view(mean, (n_batch, C), np.float32), view(rstd, (n_batch, C), np.float32)
        xh = (xv - mu[b]) * rs[b]
        s = view(sums, (n_batch, 2, C), np.float64)
        s[:] = 0
        np.add.at(s[:, 0], b, g)
        np.add.at(s[:, 1], b, g * xh)
        n = np.maximum(view(cnt, n_batch, np.int32), 1)[:, None]
        ga = view(gamma, C, np.float32) if _addr(gamma) else 1.0
        view(dx, (m, C), np.float32)[:] = ga * rs[b] * (g - (s[:, 0] / n)[b] - xh * (s[:, 1] / n)[b])
        return 0

    def spc_interp_corners(self, query, n, ts, lower, weights, stream):
        self._called("spc_interp_corners")
        ts = tuple(int(v) for v in view(ts, 3, np.int32))
        corners, _, w = R.interp_map_np(np.zeros((0, 4), np.int32), view(query, (n, 4), np.float32), ts)
        view(lower, (n, 4), np.int32)[:] = corners[:, 0]
        view(weights, (8, n), np.float32)[:] = w
        return 0

    def spc_interp_fwd(self, feats, idx, weights, n, C, K, out, stream):
        self._called("spc_interp_fwd")
        i, w = view(idx, (K, n), np.int32), view(weights, (K, n), np.float32)
        m = int(i.max()) + 1 if i.size else 0
        f = view(feats, (max(m, 1), C), np.float32)
        acc = np.zeros((n, C), np.float64)
        for k in range(K):
            ok = i[k] >= 0
            acc[ok] += w[k, ok, None].astype(np.float64) * f[i[k, ok]]
        view(out, (n, C), np.float32)[:] = acc
        return 0

    def spc_interp_bwd(self, dout, idx, weights, n, m, C, K, dfeats, stream):
        self._called("spc_interp_bwd")
        i, w = view(idx, (K, n), np.int32), view(weights, (K, n), np.float32)
        g = view(dout, (n, C), np.float32)
        acc = np.zeros((m, C), np.float64)
        for k in range(K):
            ok = i[k] >= 0
            np.add.at(acc, i[k, ok], w[k, ok, None].astype(np.float64) * g[ok])
        view(dfeats, (m, C), np.float32)[:] = acc
        return 0


def install(monkeypatch, precision: str = "fp32", prefetch_depth: int = 1) -> FakeLib:
    """Route the host layer's library calls to a FakeLib for the duration of a test."""
    from nerf_downstream_b200 import lib as L
    from nerf_downstream_b200 import ops
    from nerf_downstream_b200.me import core
    fake = FakeLib()

    def ptr(t):
        if t is None:
            return None
        if not t.is_contiguous():
            raise RuntimeError("sparseconv_b200 kernels need contiguous tensors")
        return t.data_ptr()

    monkeypatch.setattr(L, "load", lambda: fake)
    monkeypatch.setattr(L, "ptr", ptr)
    monkeypatch.setattr(L, "stream", lambda: 0)
    monkeypatch.setattr(L, "launch_count", lambda: fake.launches)
    monkeypatch.setattr(ops, "_ptr_rows", lambda x: x.data_ptr())
    monkeypatch.setattr(core, "_require_cuda", lambda dev, what: None)
    monkeypatch.setattr(core.CoordinateManager, "default_prefetch_depth", prefetch_depth)   # 4 = the product's pyramid
    monkeypatch.setattr(ops, "_default_precision", ops.PRECISIONS[precision])
    monkeypatch.setattr(ops, "_ws_bytes_cache", {})
    monkeypatch.setattr(ops, "_workspaces", {})
    return fake
