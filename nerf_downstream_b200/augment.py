"""The reference's affine-type coordinate augmentations folded into ONE affine map.

The reference augments a sample on the CPU, one numpy pass per transform (co3d_3d/src/data/transforms.py):
    RandomRotation (:339-358)               coords @ M(axis, angle)
    RandomScale (:361-373)                  coords * s
    RandomTranslation (:376-392)            coords + t
    CoordinateUniformTranslation (:284-294) coords + t
    RandomAffine (:395-427)                 coords @ (M(axis, angle) @ (diag(scale) + shear))
    RandomHorizontalFlip (:430-450)         coords[:, ax] = max(coords[:, ax]) - coords[:, ax]   (horizontal axes)
    DimensionlessCoordinates (:453-460)     coords / voxel_size
Every one of them is an affine map of the point, so a chain of them is ONE 3x3 matrix + translation — exactly the
`affine12` argument of `spc_plenoxel_decode`, which applies it while it decodes the plenoxel record on the GPU
(`pipeline.plenoxel_decode`).  `AffineChain` composes the maps in float64; the `sample_*` functions draw the random
parameters from Python's `random` and numpy's global generator IN THE REFERENCE'S ORDER, so a run seeded like the
reference (`pl.seed_everything`, train.py:251) draws the same augmentations (tests/golden/make_augment.py pins this
against the reference's classes).  The flip needs the current maximum of the coordinates along an axis — a property
of the data — which the caller supplies (`flip(axis, coord_max)`; `pipeline.plenoxel_decode_augmented` gets it from a
first decode pass).

Not affine: CoordinateDropout / RandomCrop (row subsets), CoordinateJitter / ElasticDistortion (per-point
displacements), RandomFeatureJitter — the second half of this module runs those on device tensors with the same RNG
draws, and `apply_transformations` is the reference's `Compose` over both kinds.
"""
from __future__ import annotations

import math
import random as py_random
from typing import Callable, List, Optional, Sequence

import numpy as np

AXES = {"x": 0, "y": 1, "z": 2}


def rotation_matrix(axis: Sequence[float], theta: float) -> np.ndarray:
    """`M(axis, theta) = expm(cross(eye(3), axis / |axis| * theta))` (transforms.py:334-336) in closed form (Rodrigues):
    the exponential of the skew matrix K with K[i] = e_i x u is I + sin(theta) K + (1 - cos(theta)) K^2."""
    u = np.asarray(axis, np.float64)
    u = u / np.linalg.norm(u)
    K = np.cross(np.eye(3), u)
    return np.eye(3) + math.sin(theta) * K + (1.0 - math.cos(theta)) * (K @ K)


class AffineChain:
    """c' = A c + t, built up by appending the reference's transforms in the order they are applied."""

    def __init__(self):
        self.A = np.eye(3)
        self.t = np.zeros(3)
        self.steps: List[str] = []

    # -- generic -------------------------------------------------------------------------------------------------
    def right_multiply(self, T: np.ndarray, name: str = "matrix") -> "AffineChain":
        """`coords = coords @ T` (row-vector convention of the reference)."""
        T = np.asarray(T, np.float64)
        self.A = T.T @ self.A
        self.t = T.T @ self.t
        self.steps.append(name)
        return self

    def scale(self, s) -> "AffineChain":
        s = np.asarray(s, np.float64)
        self.A = (self.A.T * s).T if s.ndim else self.A * s
        self.t = self.t * s
        self.steps.append("scale")
        return self

    def translate(self, v) -> "AffineChain":
        self.t = self.t + np.asarray(v, np.float64).reshape(3)
        self.steps.append("translate")
        return self

    def flip(self, axis: int, coord_max: float) -> "AffineChain":
        """`coords[:, axis] = coord_max - coords[:, axis]`; coord_max = the data's current maximum along the axis."""
        self.A[axis, :] = -self.A[axis, :]
        self.t[axis] = coord_max - self.t[axis]
        self.steps.append(f"flip{axis}")
        return self

    def divide(self, voxel_size: float) -> "AffineChain":
        self.A = self.A / voxel_size
        self.t = self.t / voxel_size
        self.steps.append("divide")
        return self

    # -- results -------------------------------------------------------------------------------------------------
    def apply(self, coords: np.ndarray) -> np.ndarray:
        return np.asarray(coords, np.float64) @ self.A.T + self.t

    def axis_max(self, coords: np.ndarray, axis: int) -> float:
        return float(np.max(np.asarray(coords, np.float64) @ self.A[axis] + self.t[axis]))

    def as_affine12(self) -> List[float]:
        """Row-major 3x3 then the translation: `affine12` of spc_plenoxel_decode (float32 on the device)."""
        return [float(v) for v in self.A.reshape(-1)] + [float(v) for v in self.t]

    def copy(self) -> "AffineChain":
        c = AffineChain()
        c.A, c.t, c.steps = self.A.copy(), self.t.copy(), list(self.steps)
        return c


# ---- the reference's samplers: same draws, same order -----------------------------------------------------------
def sample_random_rotation(chain: AffineChain, upright_axis: str = "z", axis_std: float = 0.01,
                           application_ratio: float = 0.9) -> bool:
    """RandomRotation.__call__ (transforms.py:352-358)."""
    if py_random.random() < application_ratio:
        axis = axis_std * np.random.randn(3)
        axis[AXES[upright_axis.lower()]] += 1
        angle = py_random.random() * 2 * np.pi
        chain.right_multiply(rotation_matrix(axis, angle), "RandomRotation")
        return True
    return False


def sample_random_scale(chain: AffineChain, scale_ratio: float = 0.1, application_ratio: float = 0.9) -> bool:
    """RandomScale.__call__ (transforms.py:368-373)."""
    if py_random.random() < application_ratio:
        chain.scale(np.random.uniform(low=1 - scale_ratio, high=1 + scale_ratio))
        chain.steps[-1] = "RandomScale"
        return True
    return False


def sample_random_translation(chain: AffineChain, max_translation: float = 3, application_ratio: float = 0.9) -> bool:
    """RandomTranslation.__call__ (transforms.py:389-392)."""
    if py_random.random() < application_ratio:
        chain.translate(2 * (np.random.rand(1, 3) - 0.5) * max_translation)
        chain.steps[-1] = "RandomTranslation"
        return True
    return False


def sample_uniform_translation(chain: AffineChain, max_translation: float = 0.2) -> bool:
    """CoordinateUniformTranslation.__call__ (transforms.py:289-294): always applied when max_translation > 0."""
    if max_translation > 0:
        chain.translate(np.random.uniform(low=-max_translation, high=max_translation, size=[1, 3]))
        chain.steps[-1] = "CoordinateUniformTranslation"
        return True
    return False


def sample_random_affine(chain: AffineChain, upright_axis: str = "z", axis_std: float = 0.1, scale_range: float = 0.2,
                         affine_range: float = 0.1, application_ratio: float = 0.9) -> bool:
    """RandomAffine.__call__ (transforms.py:416-427)."""
    if py_random.random() < application_ratio:
        axis = axis_std * np.random.randn(3)
        axis[AXES[upright_axis.lower()]] += 1
        angle = 2 * (py_random.random() - 0.5) * np.pi
        T = rotation_matrix(axis, angle) @ (np.diag(2 * (np.random.rand(3) - 0.5) * scale_range + 1)
                                            + 2 * (np.random.rand(3, 3) - 0.5) * affine_range)
        chain.right_multiply(T, "RandomAffine")
        return True
    return False


def sample_horizontal_flip(chain: AffineChain, axis_max: Callable[[AffineChain, int], float], upright_axis: str = "z",
                           application_ratio: float = 0.9) -> bool:
    """RandomHorizontalFlip.__call__ (transforms.py:445-450): both horizontal axes are mirrored about the data's
    current maximum; `axis_max(chain, axis)` returns that maximum under the chain built so far."""
    if py_random.random() < application_ratio:
        for ax in sorted(set(range(3)) - {AXES[upright_axis.lower()]}):
            chain.flip(ax, axis_max(chain, ax))
        return True
    return False


def dimensionless(chain: AffineChain, voxel_size: float = 0.02) -> bool:
    """DimensionlessCoordinates.__call__ (transforms.py:459-460)."""
    chain.divide(voxel_size)
    return True


SAMPLERS = {
    "RandomRotation": sample_random_rotation,
    "RandomScale": sample_random_scale,
    "RandomTranslation": sample_random_translation,
    "CoordinateUniformTranslation": sample_uniform_translation,
    "RandomAffine": sample_random_affine,
    "RandomHorizontalFlip": sample_horizontal_flip,
    "DimensionlessCoordinates": dimensionless,
}


def sample_chain(names: Sequence[str], axis_max: Optional[Callable[[AffineChain, int], float]] = None,
                 params: Optional[dict] = None) -> AffineChain:
    """Draw one augmentation chain for a sample: `names` in the order of `*.train_transformations`
    (scannet_plenoxel.gin:7-16, co3d_aug3.gin:3-12), per-transform keyword arguments from `params[name]` or, when
    absent, from the gin bindings of that name (`RandomRotation.upright_axis = "y"` ...).  Names that are not affine
    raise: they need the points themselves."""
    from . import ginlite
    chain = AffineChain()
    for name in names:
        fn = SAMPLERS.get(name)
        if fn is None:
            raise KeyError(f"{name} is not an affine transform (affine: {sorted(SAMPLERS)})")
        kw = dict((params or {}).get(name) or
                  {k.split(".", 1)[1]: v for k, v in ginlite.config_dict().items() if k.startswith(name + ".")})
        if name == "RandomHorizontalFlip":
            if axis_max is None:
                raise ValueError("RandomHorizontalFlip needs axis_max(chain, axis): the data's maximum along an axis")
            fn(chain, axis_max, **kw)
        else:
            fn(chain, **kw)
    return chain


# ---- the non-affine train transformations, on device tensors ------------------------------------------------------
# Same RNG draws in the same order as the reference classes (host `random` / numpy global generator), the arithmetic
# on whatever device `coords` lives on (torch ops: these are row selections and element-wise passes, not kernels of
# their own).  Together with AffineChain they cover `PlenoxelScannetDataset.train_transformations`
# (scannet_plenoxel.gin:7-16) and `Co3DDatasetBase.train_transformations` (co3d_aug3.gin:3-12).
def _index(t, idx):
    return t if t is None else t[idx]


def coordinate_dropout(coords, feats, labels, dropout_ratio: float = 0.2, application_ratio: float = 0.2):
    """CoordinateDropout.__call__ (transforms.py:256-265): keep `int(N * (1 - dropout_ratio))` random rows."""
    import torch
    if py_random.random() < application_ratio:
        n = len(coords)
        inds = np.random.choice(n, int(n * (1 - dropout_ratio)), replace=False)
        idx = torch.from_numpy(inds).to(coords.device)
        return coords[idx], _index(feats, idx), _index(labels, idx)
    return coords, feats, labels


def coordinate_jitter(coords, feats, labels, jitter_std: float = 0.5, application_ratio: float = 0.7):
    """CoordinateJitter.__call__ (transforms.py:277-281): uniform per-point noise in [-jitter_std, jitter_std)."""
    import torch
    if py_random.random() < application_ratio:
        noise = (2 * jitter_std) * (np.random.rand(len(coords), 3) - 0.5)
        coords = coords + torch.from_numpy(noise).to(device=coords.device, dtype=coords.dtype)
    return coords, feats, labels


def random_feature_jitter(coords, feats, labels, std: float = 0.01, application_ratio: float = 0.9, start_ind: int = 4,
                          feature_dim: int = 27):
    """RandomFeatureJitter.__call__ (transforms.py:35-41) — including its `randn - 0.5` (a -0.5 std shift)."""
    import torch
    if py_random.random() < application_ratio:
        noise = (np.random.randn(feats.shape[0], feature_dim) - 0.5) * std
        feats = feats.clone()
        feats[:, start_ind:start_ind + feature_dim] += torch.from_numpy(noise).to(device=feats.device, dtype=feats.dtype)
    return coords, feats, labels


def random_crop(coords, feats, labels, x, y, z, application_ratio: float = 1, max_retries: int = 10):
    """RandomCrop.__call__ (transforms.py:206-243): an axis-aligned box of size (x, y, z) at a random position inside
    the extent of the points (strict inequalities); retried while empty, the input returned when it stays empty."""
    import torch
    if py_random.random() > application_ratio:
        return coords, feats, labels
    max_size = torch.tensor([[x, y, z]], dtype=coords.dtype, device=coords.device)
    norm = coords - coords.min(0, keepdim=True).values
    coord_range = (norm.max(0, keepdim=True).values - max_size).clamp_min(0)
    if bool((coord_range == 0).all()):                      # `np.prod(coord_range == 0)`: every axis fits the box
        return coords, feats, labels
    valid, retries, sel = False, 0, None
    while not valid:
        min_box = torch.from_numpy(np.random.rand(1, 3)).to(device=coords.device, dtype=coords.dtype) * coord_range
        max_box = min_box + max_size
        sel = (norm > min_box).all(1) & (norm < max_box).all(1)
        if int(sel.sum()) > 0:
            valid = True
        retries += 1
        if retries >= max_retries:
            break
    if valid:
        return coords[sel], _index(feats, sel), _index(labels, sel)
    return coords, feats, labels


def _trilinear_grid(grid, origin, spacing, pts):
    """`scipy.interpolate.RegularGridInterpolator(axes, grid, bounds_error=0, fill_value=0)(pts)` for axes that are
    `origin + spacing * arange(n)`: linear in each axis, zero outside the grid.  grid [X, Y, Z, C], pts [N, 3]."""
    import torch
    X, Y, Z, C = grid.shape
    u = (pts - origin) / spacing
    size = torch.tensor([X - 1, Y - 1, Z - 1], dtype=pts.dtype, device=pts.device)
    inside = ((u >= 0) & (u <= size)).all(1)
    i0 = torch.minimum(torch.floor(u).clamp_min(0), size - 1).long()
    f = (u - i0.to(pts.dtype)).clamp(0, 1)
    flat = grid.reshape(-1, C)
    out = torch.zeros((pts.shape[0], C), dtype=pts.dtype, device=pts.device)
    for k in range(8):
        b = [(k >> a) & 1 for a in range(3)]
        w = torch.ones(pts.shape[0], dtype=pts.dtype, device=pts.device)
        for a in range(3):
            w = w * (f[:, a] if b[a] else 1 - f[:, a])
        lin = ((i0[:, 0] + b[0]) * Y + (i0[:, 1] + b[1])) * Z + (i0[:, 2] + b[2])
        out = out + w[:, None] * flat[lin]
    return torch.where(inside[:, None], out, torch.zeros((), dtype=pts.dtype, device=pts.device))


def elastic_distortion(coords, feats, labels, distortion_params=((4, 16), (8, 24)), application_ratio: float = 0.9):
    """ElasticDistortion.__call__ (transforms.py:544-596): per (granularity, magnitude) a Gaussian noise grid over the
    extent of the points (drawn and box-blurred on the host — it is tiny), trilinearly interpolated at every point on
    the device and added as a displacement."""
    import scipy.ndimage
    import torch
    if distortion_params is None or not (py_random.random() < application_ratio):
        return coords, feats, labels
    blurs = [np.ones(s, np.float32) / 3 for s in ((3, 1, 1, 1), (1, 3, 1, 1), (1, 1, 3, 1))]
    for granularity, magnitude in distortion_params:
        cmin = coords.min(0).values
        noise_dim = (torch.div((coords - cmin).max(0).values, granularity, rounding_mode="floor")).long().cpu().numpy() + 3
        noise = np.random.randn(*noise_dim, 3).astype(np.float32)
        for _ in range(2):
            for blur in blurs:
                noise = scipy.ndimage.convolve(noise, blur, mode="constant", cval=0)
        # axes: linspace(cmin - g, cmin + g * (dim - 2), dim) == cmin - g + g * arange(dim)
        grid = torch.from_numpy(noise).to(device=coords.device, dtype=coords.dtype)
        coords = coords + _trilinear_grid(grid, cmin - granularity, float(granularity), coords) * magnitude
    return coords, feats, labels


POINT_TRANSFORMS = {
    "CoordinateDropout": coordinate_dropout,
    "CoordinateJitter": coordinate_jitter,
    "RandomFeatureJitter": random_feature_jitter,
    "RandomCrop": random_crop,
    "ElasticDistortion": elastic_distortion,
}


def apply_transformations(names: Sequence[str], coords, feats=None, labels=None, params: Optional[dict] = None):
    """`Compose([...])(coords, feats, labels)` (transforms.py:710-719) for a list of transform names in dataset order.
    Runs of affine transforms are folded into ONE matrix + translation and applied in a single pass when a
    point-wise transform (or the end of the list) needs the actual coordinates.  Parameters per transform from
    `params[name]` or the gin bindings of that name.  coords [N, 3] torch tensor (any device), float32 or float64."""
    import torch

    from . import ginlite

    def kwargs(name):
        return dict((params or {}).get(name) or
                    {k.split(".", 1)[1]: v for k, v in ginlite.config_dict().items() if k.startswith(name + ".")})

    chain = AffineChain()

    def flush(c):
        nonlocal chain
        if chain.steps:
            A = torch.tensor(chain.A, dtype=c.dtype, device=c.device)
            t = torch.tensor(chain.t, dtype=c.dtype, device=c.device)
            c = c @ A.T + t
            chain = AffineChain()
        return c

    def axis_max(ch, axis):
        a = torch.tensor(ch.A[axis], dtype=torch.float64, device=coords.device)
        return float((coords.double() @ a).max()) + float(ch.t[axis])

    for name in names:
        if name in SAMPLERS:
            if name == "RandomHorizontalFlip":
                SAMPLERS[name](chain, axis_max, **kwargs(name))
            else:
                SAMPLERS[name](chain, **kwargs(name))
        elif name in POINT_TRANSFORMS:
            coords = flush(coords)
            coords, feats, labels = POINT_TRANSFORMS[name](coords, feats, labels, **kwargs(name))
        else:
            raise KeyError(f"transformation {name!r} is not built (have {sorted(list(SAMPLERS) + list(POINT_TRANSFORMS))})")
    return flush(coords), feats, labels
