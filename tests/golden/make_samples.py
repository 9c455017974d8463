"""Generate tests/golden/samples_ref.npz with the REFERENCE's own Dataset.__getitem__ code,
    co3d_3d/src/data/co3d.py    : Co3DDatasetBase.__getitem__ (:178-235)
    co3d_3d/src/data/scannet.py : PlenoxelScannetDataset.load_data / __getitem__ (:558-654)
imported unchanged from /root/reference.  The classes are instantiated without their constructors (those read file
lists from disk) and `load_data` / the npz reader is replaced by a synthetic in-memory plenoxel record, so that what
runs is exactly the reference's record -> sample arithmetic: (i,j,k) decode, SH dequantisation, void labelling, lattice
thinning, normalisation, train transformations (its own Compose, parameters injected through `ginlite`), feature
selection, label map.  `matplotlib` / `plyfile` (imported by data/utils.py for plotting / .ply reading, not installed
here) are stubbed with empty modules.

Run from the repository root:  python tests/golden/make_samples.py
"""
import importlib
import random
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
from nerf_downstream_b200 import ginlite  # noqa: E402

GIN = """
RandomRotation.upright_axis = "y"
RandomHorizontalFlip.upright_axis = "y"
RandomAffine.upright_axis = "y"
RandomAffine.application_ratio = 0.7
CoordinateDropout.application_ratio = 0.9
CoordinateUniformTranslation.max_translation = 0.2
RandomScale.scale_ratio = 0.40
RandomFeatureJitter.start_ind = 4
RandomFeatureJitter.feature_dim = 27
RandomCrop.x = 150
RandomCrop.y = 150
RandomCrop.z = 150
ElasticDistortion.distortion_params = [(4, 16)]
ElasticDistortion.application_ratio = 0.7
"""
CO3D_TRANSFORMS = ["RandomRotation", "RandomAffine", "CoordinateDropout", "RandomHorizontalFlip",
                   "CoordinateUniformTranslation", "CoordinateJitter", "RandomScale", "RandomFeatureJitter"]
SCANNET_TRANSFORMS = ["RandomRotation", "RandomCrop", "RandomAffine", "CoordinateDropout", "RandomHorizontalFlip",
                      "RandomTranslation", "ElasticDistortion"]


def load_reference():
    sys.modules["gin"] = ginlite
    for name in ("matplotlib", "matplotlib.cm", "plyfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["plyfile"].PlyData = object
    import MinkowskiEngine  # noqa: F401  (this repository's package)
    for name in ("co3d_3d", "co3d_3d.src", "co3d_3d.src.data"):
        mod = types.ModuleType(name)
        mod.__path__ = [str(REF / name.replace(".", "/"))]
        sys.modules[name] = mod
    return (importlib.import_module("co3d_3d.src.data.co3d"), importlib.import_module("co3d_3d.src.data.scannet"),
            importlib.import_module("co3d_3d.src.data.transforms"))


def record(rng, reso, n):
    links = np.sort(rng.choice(reso ** 3, size=n, replace=False)).astype(np.int64)
    return dict(links=links, density=rng.standard_normal((n, 1)).astype(np.float32),
                sh=rng.integers(0, 256, (n, 27), dtype=np.uint8), sh_scale=np.float32(2.0 / 255.0), sh_min=np.float32(-1.0),
                reso=np.array([reso] * 3), labels=rng.integers(0, 41, n).astype(np.int64),
                dists=rng.uniform(0, 0.08, n).astype(np.float32))


def main():
    co3d, scannet, T = load_reference()
    ginlite.clear_config()
    ginlite.parse_config(GIN)
    rng = np.random.default_rng(20261019)
    out = {}
    rec = record(rng, 64, 4000)
    for k, v in rec.items():
        out[f"record/{k}"] = np.array(v)                 # a copy: scannet.py normalises the density IN PLACE

    # ---- CO3D -----------------------------------------------------------------------------------------------
    # (no transformation case here: co3d.py hands torch tensors to the numpy-based transforms, which this image's
    #  numpy / torch versions reject — `np.max(tensor)`; the transformation lists themselves are pinned on numpy inputs
    #  by make_augment.py, and the ScanNet path below converts to numpy first and runs them)
    for case, (feats, names, seed) in {"co3d_plain": (["sh"], [], 0),
                                       "co3d_xyz_density": (["xyzs", "density", "ones"], [], 0)}.items():
        ds = object.__new__(co3d.Co3DDatasetBase)
        ds.files, ds.CLASS_LABELS, ds.features = [("car", "x")], ["apple", "car"], feats
        ds.transformations = T.Compose([T.__dict__[t]() for t in names]) if names else None
        sh = rec["sh"].astype(np.float32) * rec["sh_scale"] + rec["sh_min"]
        ds.load_data = lambda inst: dict(links=torch.from_numpy(rec["links"].copy()), density=torch.from_numpy(rec["density"].copy()),
                                         sh=torch.from_numpy(sh), reso=[64, 64, 64])
        random.seed(seed)
        np.random.seed(seed)
        s = ds[0]
        out[f"{case}/coordinates"] = np.asarray(s["coordinates"], np.float32)      # the reference returns float32
        out[f"{case}/features"] = np.asarray(s["features"], np.float32)
        out[f"{case}/xyzs"] = np.asarray(s["xyzs"], np.float32)
        print(case, out[f"{case}/coordinates"].shape, out[f"{case}/features"].shape, int(s["labels"][0]))

    # ---- ScanNet plenoxel ---------------------------------------------------------------------------------------
    for case, (feats, names, seed, void, ignore_thres) in {
            "scannet_plain": (["sh"], [], 0, None, None), "scannet_void": (["density", "sh"], [], 0, 40, None),
            "scannet_aug": (["sh"], SCANNET_TRANSFORMS, 2, None, None)}.items():
        ds = object.__new__(scannet.PlenoxelScannetDataset)
        ignore_label = -255
        ds.files, ds.features, ds.voxel_size = ["scene0"], feats, 0.02
        ds.ignore_label, ds.void_label = ignore_label, (void if void is not None else ignore_label)
        ds.valid_thres, ds.ignore_thres, ds.downsample_mode, ds.downsample_stride = 0.05, ignore_thres, 1, 2
        ds.scene_scales = {"scene0": 0.34}
        ds.data_root = "/nonexistent"
        ds.transformations = T.Compose([T.__dict__[t]() for t in names]) if names else None
        label_map, n_used = dict(), 0
        for lab in range(ds.NUM_LABELS):                      # the constructor's label map (scannet.py:518-528)
            if lab in ds.IGNORE_LABELS:
                label_map[lab] = ignore_label
            else:
                label_map[lab] = n_used
                n_used += 1
        label_map[ignore_label] = ignore_label
        if void is not None and void != ignore_label:
            label_map[void] = n_used
        ds.label_map = label_map
        real_load = np.load
        np.load = lambda path: {k: np.array(rec[k]) for k in ("links", "density", "sh", "sh_scale", "sh_min", "reso", "labels", "dists")}
        try:
            random.seed(seed)
            np.random.seed(seed)
            s = ds[0]
        finally:
            np.load = real_load
        for k in ("coordinates", "features", "labels", "dists"):
            out[f"{case}/{k}"] = np.asarray(s[k], np.float32 if k != "labels" else np.int64)
        print(case, out[f"{case}/coordinates"].shape, out[f"{case}/features"].shape, np.unique(out[f"{case}/labels"])[:6])
    path = Path(__file__).with_name("samples_ref.npz")
    np.savez_compressed(path, **out)
    Path(__file__).with_name("samples_ref.gin").write_text(GIN)
    print("wrote", path)


if __name__ == "__main__":
    main()
