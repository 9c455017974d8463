"""Golden vectors produced by the REFERENCE's own in-tree conv restatement
(co3d_3d/src/models/mink/modules/sparse_conv.py:57-152,160-264; see tests/golden/make_golden.py).
CPU: the oracle reproduces them.  GPU: the CUDA kernels reproduce them through the C ABI."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_ops as R

GOLD = np.load(Path(__file__).parent / "golden" / "conv_forward_ref.npz")
CASES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_restatement(name):
    nbr, feats, w, want = (GOLD[f"{name}/{k}"] for k in ("nbr", "feats", "w", "out"))
    got = R.conv_forward(torch.from_numpy(feats), torch.from_numpy(w), nbr).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    got64 = R.conv_forward(torch.from_numpy(feats).double(), torch.from_numpy(w).double(), nbr).numpy()
    np.testing.assert_allclose(got64, want, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference_restatement(cuda_device, name):
    from nerf_downstream_b200 import lib as L
    from nerf_downstream_b200 import ops
    nbr, feats, w, want = (GOLD[f"{name}/{k}"] for k in ("nbr", "feats", "w", "out"))
    K, m_out = nbr.shape
    km = ops.KernelMap(torch.from_numpy(nbr).to(cuda_device), None, K, feats.shape[0], m_out)
    out = ops.conv_fwd_raw(torch.from_numpy(feats).to(cuda_device), torch.from_numpy(w).to(cuda_device), None, km,
                           L.PREC_FP32)
    err = np.abs(out.cpu().numpy() - want)
    assert (err <= 1e-4 * (1 + np.abs(want))).all(), err.max()
