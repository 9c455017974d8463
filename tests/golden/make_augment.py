"""Generate tests/golden/augment_ref.json with the REFERENCE's own augmentation classes,
    co3d_3d/src/data/transforms.py : RandomRotation (:339), RandomScale (:361), RandomTranslation (:376),
    CoordinateUniformTranslation (:284), RandomAffine (:395), RandomHorizontalFlip (:430), DimensionlessCoordinates (:453),
imported unchanged from /root/reference (on `ginlite` for gin and this repository's `MinkowskiEngine` package, which
the file imports but these classes do not use).  For every seed, Python's `random` and numpy's global generator are
seeded like `pl.seed_everything` does, the transforms are applied one after the other to the same 40 points, and the
resulting coordinates are stored; `nerf_downstream_b200.augment` must reach them with ONE composed affine map.

Run from the repository root:  python tests/golden/make_augment.py
"""
import importlib.util
import json
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from nerf_downstream_b200 import ginlite  # noqa: E402

SEQUENCES = {
    # name -> list of (class name, constructor kwargs), in application order
    "scannet_like": [("RandomRotation", dict(upright_axis="y", application_ratio=0.7)),
                     ("RandomAffine", dict(upright_axis="y", application_ratio=0.7)),
                     ("RandomHorizontalFlip", dict(upright_axis="y", application_ratio=0.7)),
                     ("RandomTranslation", dict(max_translation=3, application_ratio=0.7))],
    "co3d_like": [("RandomRotation", dict(upright_axis="y")),
                  ("RandomAffine", dict(upright_axis="y")),
                  ("RandomHorizontalFlip", dict(upright_axis="y")),
                  ("CoordinateUniformTranslation", dict(max_translation=0.2)),
                  ("RandomScale", dict(scale_ratio=0.4)),
                  ("DimensionlessCoordinates", dict(voxel_size=0.02))],
    "z_up": [("RandomHorizontalFlip", dict(upright_axis="z", application_ratio=1.0)),
             ("RandomRotation", dict(upright_axis="z", axis_std=0.2, application_ratio=1.0)),
             ("RandomScale", dict(scale_ratio=0.1, application_ratio=0.5))],
}


PIPELINES = {
    # PlenoxelScannetDataset.train_transformations (scannet_plenoxel.gin:7-16) with its gin parameters
    "scannet_plenoxel": [("RandomRotation", dict(upright_axis="y")),
                         ("RandomCrop", dict(x=60, y=60, z=60)),
                         ("RandomAffine", dict(upright_axis="y", application_ratio=0.7)),
                         ("CoordinateDropout", dict()),
                         ("RandomFeatureJitter", dict(start_ind=4, feature_dim=27)),
                         ("RandomHorizontalFlip", dict(upright_axis="y")),
                         ("RandomTranslation", dict()),
                         ("ElasticDistortion", dict(distortion_params=[(4, 16)], application_ratio=0.7))],
    # Co3DDatasetBase.train_transformations (co3d_aug3.gin:3-12)
    "co3d_aug3": [("RandomRotation", dict(upright_axis="y")),
                  ("RandomAffine", dict(upright_axis="y")),
                  ("CoordinateDropout", dict(application_ratio=0.9)),
                  ("RandomHorizontalFlip", dict(upright_axis="y")),
                  ("CoordinateUniformTranslation", dict(max_translation=0.2)),
                  ("CoordinateJitter", dict()),
                  ("RandomScale", dict(scale_ratio=0.4)),
                  ("RandomFeatureJitter", dict(start_ind=4, feature_dim=27))],
}


def make_pipelines(T):
    """Full transformation lists (affine AND point-wise) through the reference's Compose -> augment_pipeline_ref.npz."""
    rng = np.random.default_rng(20261018)
    n = 400
    points = rng.uniform(-40, 90, (n, 3))
    feats = rng.standard_normal((n, 31))
    labels = rng.integers(0, 20, n)
    out = {"points": points, "feats": feats, "labels": labels}
    for name, seq in PIPELINES.items():
        for seed in range(4):
            random.seed(seed)
            np.random.seed(seed)
            comp = T.Compose([getattr(T, cls)(**kw) for cls, kw in seq])
            c, f, l = comp(points.copy(), feats.copy(), labels.copy())
            out[f"{name}/{seed}/coords"] = np.asarray(c, np.float64)
            out[f"{name}/{seed}/feats"] = np.asarray(f, np.float64)
            out[f"{name}/{seed}/labels"] = np.asarray(l, np.int64)
            print(name, seed, c.shape)
    path = Path(__file__).with_name("augment_pipeline_ref.npz")
    np.savez_compressed(path, **out)
    (Path(__file__).with_name("augment_pipeline_ref.json")).write_text(
        json.dumps({k: [[c, kw] for c, kw in v] for k, v in PIPELINES.items()}))
    print("wrote", path)


def main():
    sys.modules["gin"] = ginlite
    import MinkowskiEngine  # noqa: F401  (this repository's drop-in package; imported by the reference file)
    spec = importlib.util.spec_from_file_location("ref_transforms", "/root/reference/co3d_3d/src/data/transforms.py")
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    points = np.random.default_rng(20261017).uniform(-40, 90, (40, 3))
    out = {"points": points.tolist(), "cases": []}
    for seq_name, seq in SEQUENCES.items():
        for seed in range(6):
            random.seed(seed)
            np.random.seed(seed)
            transforms = [getattr(T, name)(**kw) for name, kw in seq]
            coords = points.copy()
            for t in transforms:
                coords, _, _ = t(coords, None, None)
                coords = np.array(coords, dtype=np.float64)            # some transforms work in place
            out["cases"].append({"sequence": seq_name, "seed": seed, "coords": coords.tolist()})
    out["sequences"] = {k: [[n, kw] for n, kw in v] for k, v in SEQUENCES.items()}
    path = Path(__file__).with_name("augment_ref.json")
    path.write_text(json.dumps(out))
    print("wrote", path, len(out["cases"]), "cases")
    make_pipelines(T)


if __name__ == "__main__":
    main()
