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

Not affine, hence not here: CoordinateDropout / RandomCrop (row subsets), CoordinateJitter / ElasticDistortion
(per-point noise), the feature jitters.
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
