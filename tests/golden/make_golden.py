"""Generate tests/golden/conv_forward_ref.npz by running the REFERENCE's own in-tree restatement of
the sparse-convolution forward pass,
    WeightSparseConvolutionFunction.forward / WeightSparseConvolutionTransposeFunction.forward
    (/root/reference/co3d_3d/src/models/mink/modules/sparse_conv.py:57-152, :160-264),
imported unchanged in the build container, on kernel maps given in ME's documented layout
{k: IntTensor[2, n_k]} (row 0 = in rows, row 1 = out rows).  The maps are produced by
oracle/ref_ops.py, so the fixture pins the ARITHMETIC given a map (gather / per-offset product /
scatter-add / offset order / zero-initialised output), not the map itself.

Run from the repository root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_ops as R  # noqa: E402
from tests import ref_harness  # noqa: E402


class FakeManager:
    """Stands in for ME's CoordinateManager: only size() and kernel_map() are called
    (sparse_conv.py:80, :90-96, :197-204)."""

    def __init__(self, n_out, pairs):
        self.n_out = n_out
        self.pairs = {k: torch.from_numpy(v.copy()).int() for k, v in pairs.items()}

    def size(self, key):
        return self.n_out

    def kernel_map(self, in_key, out_key, stride, kernel_size, dilation, is_transpose=False):
        return self.pairs


class FakeGenerator:
    def __init__(self, ks, stride):
        self.kernel_size, self.kernel_stride, self.kernel_dilation = list(ks), [stride] * 3, [1, 1, 1]


def main():
    sc = ref_harness.load("co3d_3d.src.models.mink.modules.sparse_conv")
    rng = np.random.default_rng(20261017)
    out = {}
    cases = [("k3s1", (3, 3, 3), 1, 5, 7, False), ("k2s2", (2, 2, 2), 2, 8, 6, False),
             ("k3s2", (3, 3, 3), 2, 4, 9, False), ("k2s2_tr", (2, 2, 2), 2, 6, 5, True),
             ("k1s2", (1, 1, 1), 2, 3, 4, False)]
    for name, ks, stride, cin, cout, transpose in cases:
        n = 400
        c = np.empty((n, 4), np.float32)
        c[:, 0] = rng.integers(0, 2, n)
        c[:, 1:] = rng.uniform(-5, 5, (n, 3))
        mgr = R.OracleManager(c)
        ts_out = mgr.stride((1, 1, 1), (stride,) * 3)
        if transpose:
            nbr = mgr.kernel_map(ts_out, (1, 1, 1), ks, transpose=True)   # in = coarse, out = fine
        else:
            nbr = mgr.kernel_map((1, 1, 1), ts_out, ks)
        m_in = mgr.maps[ts_out if transpose else (1, 1, 1)].shape[0]
        K, m_out = nbr.shape
        feats = rng.standard_normal((m_in, cin)).astype(np.float32)
        w = (rng.standard_normal((K, cin, cout)) / np.sqrt(K * cin)).astype(np.float32)   # ME layout (K,Cin,Cout)
        pairs = R.pairs_from_dense(nbr)
        fn = sc.WeightSparseConvolutionTransposeFunction if transpose else sc.WeightSparseConvolutionFunction
        # the reference multiplies W_k @ F^T with W_k stored (Cout, Cin): hand it the transposed kernels
        w_list = [torch.from_numpy(w[k].T.copy()) for k in range(K)]
        y, _ = fn.apply(torch.from_numpy(feats), w_list, FakeGenerator(ks, stride), None, sc.CoordinateMapKey(4), sc.CoordinateMapKey(4),
                        FakeManager(m_out, pairs), sc.SparseConvMode.SPARSE, sorted(pairs))
        out[f"{name}/nbr"] = nbr
        out[f"{name}/feats"] = feats
        out[f"{name}/w"] = w
        out[f"{name}/out"] = y.numpy()
        print(name, "pairs", int((nbr >= 0).sum()), "out", y.shape)
    np.savez_compressed(Path(__file__).with_name("conv_forward_ref.npz"), **out)


if __name__ == "__main__":
    main()
