"""Affine-type augmentations folded into one map: `augment.AffineChain` against the reference's
own transform classes applied one after the other (tests/golden/augment_ref.json, made by tests/golden/make_augment.py
importing co3d_3d/src/data/transforms.py unchanged) under identical `random` / numpy seeds."""
import json
import random
from pathlib import Path

import numpy as np
import pytest

from nerf_downstream_b200 import augment, ginlite

GOLDEN = Path(__file__).resolve().parent / "golden" / "augment_ref.json"


def test_rotation_matrix_is_the_skew_exponential():
    from scipy.linalg import expm, norm
    rng = np.random.default_rng(0)
    for _ in range(20):
        axis, theta = rng.standard_normal(3), rng.uniform(-2 * np.pi, 2 * np.pi)
        ref = expm(np.cross(np.eye(3), axis / norm(axis) * theta))              # transforms.py:334-336
        got = augment.rotation_matrix(axis, theta)
        assert np.abs(got - ref).max() < 1e-12 and abs(np.linalg.det(got) - 1) < 1e-12


def test_chain_reproduces_reference_transform_sequences():
    g = json.loads(GOLDEN.read_text())
    pts = np.array(g["points"])
    assert len(g["cases"]) == 18
    applied = set()
    for case in g["cases"]:
        seq = g["sequences"][case["sequence"]]
        random.seed(case["seed"])
        np.random.seed(case["seed"])
        chain = augment.sample_chain([n for n, _ in seq], axis_max=lambda ch, ax: ch.axis_max(pts, ax),
                                     params={n: kw for n, kw in seq})
        want = np.array(case["coords"])
        err = np.abs(chain.apply(pts) - want).max()
        assert err <= 1e-12 * max(1.0, np.abs(want).max()) + 1e-12, (case["sequence"], case["seed"], err)
        applied.update(chain.steps)
        # the 12 floats the kernel gets describe the same map
        a = np.array(chain.as_affine12())
        assert np.allclose(pts @ a[:9].reshape(3, 3).T + a[9:], want, rtol=1e-12, atol=1e-9)
    assert {"RandomRotation", "RandomAffine", "RandomScale", "RandomTranslation", "CoordinateUniformTranslation",
            "flip0", "flip1", "flip2", "divide"} <= applied


def test_chain_parameters_come_from_gin_bindings():
    ginlite.clear_config()
    try:
        ginlite.parse_config('RandomRotation.upright_axis = "y"\nRandomRotation.application_ratio = 1.0\n'
                             'RandomRotation.axis_std = 0.0\nRandomScale.scale_ratio = 0.4\n'
                             'RandomScale.application_ratio = 1.0')
        random.seed(3)
        np.random.seed(3)
        chain = augment.sample_chain(["RandomRotation", "RandomScale"])
        assert chain.steps == ["RandomRotation", "RandomScale"]
        s = np.cbrt(np.linalg.det(chain.A))
        assert 0.6 <= s <= 1.4
        R = chain.A / s
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.allclose(R[1], [0, 1, 0], atol=1e-12)   # about y
        assert np.allclose(chain.t, 0)
    finally:
        ginlite.clear_config()
    with pytest.raises(KeyError, match="not an affine transform"):
        augment.sample_chain(["ElasticDistortion"])
    with pytest.raises(ValueError, match="axis_max"):
        augment.sample_chain(["RandomHorizontalFlip"], params={"RandomHorizontalFlip": {"application_ratio": 1.0}})


def test_flip_and_composition_algebra():
    pts = np.random.default_rng(1).uniform(-5, 5, (50, 3))
    c = augment.AffineChain().scale(2.0).translate([1, 2, 3])
    c.flip(0, c.axis_max(pts, 0))
    got = c.apply(pts)
    want = pts * 2.0 + [1, 2, 3]
    want[:, 0] = want[:, 0].max() - want[:, 0]
    assert np.allclose(got, want) and got[:, 0].min() == pytest.approx(0.0, abs=1e-12)
    d = c.copy().divide(0.02).scale([1.0, 2.0, 3.0])
    assert np.allclose(d.apply(pts), want / 0.02 * [1.0, 2.0, 3.0]) and np.allclose(c.apply(pts), want)


def test_full_transformation_lists_match_reference_compose():
    """Affine AND point-wise transforms through `augment.apply_transformations` (affine runs folded into one matrix,
    point-wise ones on the tensors) == the reference's `Compose` of its own classes under identical seeds: same rows
    survive crop / dropout in the same order, coordinates and jittered features agree to rounding."""
    import torch
    ref = np.load(GOLDEN.with_name("augment_pipeline_ref.npz"))
    seqs = json.loads(GOLDEN.with_name("augment_pipeline_ref.json").read_text())
    pts, feats, labels = ref["points"], ref["feats"], ref["labels"]
    seen_rows = set()
    for name, seq in seqs.items():
        for seed in range(4):
            random.seed(seed)
            np.random.seed(seed)
            c, f, l = augment.apply_transformations([n for n, _ in seq], torch.from_numpy(pts.copy()),
                                                    torch.from_numpy(feats.copy()), torch.from_numpy(labels.copy()),
                                                    params={n: kw for n, kw in seq})
            wc, wf, wl = ref[f"{name}/{seed}/coords"], ref[f"{name}/{seed}/feats"], ref[f"{name}/{seed}/labels"]
            assert c.shape == wc.shape, (name, seed, c.shape, wc.shape)
            assert (l.numpy() == wl).all()                                   # identical rows, identical order
            assert np.abs(c.numpy() - wc).max() <= 1e-9 * max(1.0, np.abs(wc).max()), (name, seed)
            assert np.abs(f.numpy() - wf).max() <= 1e-12
            seen_rows.add(c.shape[0])
    assert len(seen_rows) > 2 and min(seen_rows) < 400                       # crops / dropouts really happened


def test_apply_transformations_float32_and_errors():
    import torch
    rng = np.random.default_rng(2)
    pts = torch.from_numpy(rng.uniform(0, 50, (500, 3)).astype(np.float32))
    random.seed(1)
    np.random.seed(1)
    c, f, l = augment.apply_transformations(["RandomRotation", "CoordinateJitter", "RandomScale"], pts,
                                            params={"RandomRotation": dict(upright_axis="y", application_ratio=1.0),
                                                    "CoordinateJitter": dict(application_ratio=1.0),
                                                    "RandomScale": dict(application_ratio=1.0)})
    assert c.dtype == torch.float32 and c.shape == (500, 3) and f is None and l is None
    with pytest.raises(KeyError, match="not built"):
        augment.apply_transformations(["ChromaticJitter"], pts)
    # a box larger than the cloud leaves it untouched (transforms.py:216-218)
    same, _, _ = augment.random_crop(pts, None, None, 1000, 1000, 1000)
    assert same is pts
