"""Row selections of a plenoxel record on the GPU: RandomCrop / CoordinateDropout as a row list and one
decode of the kept records — bit-identical to the CPU restatement (oracle/ref_ops.py::random_crop_select_np, pinned on
the reference's RandomCrop through augment.random_crop / tests/golden) and to decoding everything first."""
import random as py_random

import numpy as np
import pytest
import torch

from nerf_downstream_b200 import augment, pipeline
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _record(seed, n, reso, dev):
    rng = np.random.RandomState(seed)
    links = np.sort(rng.choice(reso[0] * reso[1] * reso[2], n, replace=False)).astype(np.int32)
    sh = rng.randint(0, 256, size=(n, 27)).astype(np.uint8)
    return links, sh, torch.from_numpy(links).to(dev), torch.from_numpy(sh).to(dev)


@pytest.mark.parametrize("n,affine", [(300_000, None), (1_000_003, [0.9, 0.1, 0, -0.1, 0.9, 0.02, 0, -0.02, 1.1, 3.5, -2.0, 0.25]),
                                      (777, None)])
def test_crop_select_equals_the_oracle(cuda_device, n, affine):
    reso = (256, 200, 180)
    links, sh, links_d, sh_d = _record(n, n, reso, cuda_device)
    coords, _ = R.plenoxel_decode_np(links, np.zeros((n, 0), np.uint8), 1.0, 0.0, reso, affine=affine)
    for u3, size3 in (((0.3, 0.6, 0.1), (100.0, 90.0, 400.0)), ((0.0, 0.999, 0.5), (17.0, 33.0, 21.0)),
                      ((0.5, 0.5, 0.5), (1000.0, 1000.0, 1000.0))):
        keep, fits = R.random_crop_select_np(coords[:, 1:], u3, size3)
        rows, fits_d = pipeline.plenoxel_crop_rows(links_d, reso, size3, u3, None, affine)
        assert fits_d == fits
        assert np.array_equal(rows.cpu().numpy(), keep)
    # on a previous selection (rows in arbitrary order, as CoordinateDropout leaves them)
    sel = np.random.RandomState(1).choice(n, n // 2, replace=False).astype(np.int32)
    keep, fits = R.random_crop_select_np(coords[sel, 1:], (0.2, 0.4, 0.6), (120.0, 80.0, 60.0))
    rows, fits_d = pipeline.plenoxel_crop_rows(links_d, reso, (120.0, 80.0, 60.0), (0.2, 0.4, 0.6),
                                               torch.from_numpy(sel).to(cuda_device), affine)
    assert fits_d == fits and np.array_equal(rows.cpu().numpy(), sel[keep])


def test_decode_of_a_row_list_equals_rows_of_the_full_decode(cuda_device):
    reso = (128, 128, 128)
    n = 200_001
    links, sh, links_d, sh_d = _record(5, n, reso, cuda_device)
    aff = [1.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0, 0.5, 0.25, -3.0]
    rows = torch.from_numpy(np.random.RandomState(2).choice(n, 123_457, replace=False).astype(np.int32)).to(cuda_device)
    c_all, f_all = pipeline.plenoxel_decode(links_d, sh_d, 2.0 / 255, -1.0, reso, batch_index=3, affine=aff)
    c, f = pipeline.plenoxel_decode(links_d, sh_d, 2.0 / 255, -1.0, reso, batch_index=3, affine=aff, rows=rows)
    assert torch.equal(c, c_all[rows.long()]) and torch.equal(f, f_all[rows.long()])
    # a feature width that is not a multiple of four takes the scalar path
    c2, f2 = pipeline.plenoxel_decode(links_d, sh_d[:, :9].contiguous(), 0.5, 0.0, reso, rows=rows)
    assert torch.equal(f2, (sh_d[:, :9].float() * 0.5)[rows.long()])


def test_select_rows_equals_the_transforms_on_decoded_tensors(cuda_device):
    reso = (256, 256, 256)
    links, sh, links_d, sh_d = _record(9, 400_000, reso, cuda_device)
    steps = [("RandomCrop", dict(x=150, y=120, z=300, application_ratio=1.0)),
             ("CoordinateDropout", dict(dropout_ratio=0.2, application_ratio=1.0)),
             ("RandomCrop", dict(x=90, y=300, z=300, application_ratio=1.0))]
    py_random.seed(11)
    np.random.seed(11)
    rows = pipeline.plenoxel_select_rows(links_d, reso, steps)
    c_a, f_a = pipeline.plenoxel_decode(links_d, sh_d, 2.0 / 255, -1.0, reso, rows=rows)
    py_random.seed(11)
    np.random.seed(11)
    c_all, f_all = pipeline.plenoxel_decode(links_d, sh_d, 2.0 / 255, -1.0, reso)
    xyz, f_b, _ = augment.random_crop(c_all[:, 1:], f_all, None, 150, 120, 300, 1.0)
    xyz, f_b, _ = augment.coordinate_dropout(xyz, f_b, None, 0.2, 1.0)
    xyz, f_b, _ = augment.random_crop(xyz, f_b, None, 90, 300, 300, 1.0)
    assert 0 < xyz.shape[0] < 400_000
    assert torch.equal(c_a[:, 1:], xyz) and torch.equal(f_a, f_b)
