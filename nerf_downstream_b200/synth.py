"""Synthetic plenoxel-shaped inputs for the BASELINE configs.

No dataset is reachable (no network), so the benchmark and the parity tests draw seeded synthetic
voxel sets with the SHAPE of the reference's inputs:
  * config 1/3 — CO3D objects: integer lattice in 128^3 (co3d_3d/src/data/co3d.py:171), emulated
    augmentation rotate-about-y / scale U[0.6,1.4] / translate (configs/co3d_aug3.gin:16-21), float
    coordinates with duplicates after floor, 27 SH channels dequantised as u8*scale+min (co3d.py:169);
  * config 2 — ScanNet rooms at 2 cm voxels (configs/scannet_plenoxel.gin:24): walls, floor, ceiling,
    boxes and volumetric "haze", jittered float coordinates (data/scannet.py:612-615).
Everything is numpy on the host; tensors are handed to the engine the way `collate_mink`
(co3d_3d/src/data/utils.py:25-50) does: float32 [N,4] coordinates with the batch index in column 0.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def sh_features(rng: np.random.Generator, n: int, channels: int = 27) -> np.ndarray:
    """uint8 codes dequantised like the reference: u8 * scale + min with scale=2/255, min=-1."""
    u8 = rng.integers(0, 256, size=(n, channels), dtype=np.uint8)
    return (u8.astype(np.float32) * np.float32(2.0 / 255.0) - np.float32(1.0)).astype(np.float32)


def _box_shell(occ: np.ndarray, lo, hi, t: int):
    lo = [max(0, int(v)) for v in lo]
    hi = [min(s, int(v)) for v, s in zip(hi, occ.shape)]
    if any(h - l <= 0 for l, h in zip(lo, hi)):
        return
    sl = tuple(slice(l, h) for l, h in zip(lo, hi))
    box = occ[sl]
    inner = tuple(slice(t, max(t, s - t)) for s in box.shape)
    keep = box[inner].copy()
    box[...] = True
    box[inner] = keep


def room_scene(rng: np.random.Generator, target_voxels: int = 1_000_000, base=(300, 130, 300)) -> np.ndarray:
    """Occupied integer voxels [M,3] of one ScanNet-shaped room, M == target_voxels (+-0)."""
    s = float(np.sqrt(target_voxels / 1.0e6))
    dims = tuple(max(12, int(round(b * s))) for b in base)
    occ = np.zeros(dims, dtype=bool)
    _box_shell(occ, (0, 0, 0), dims, 3)  # floor, ceiling, four walls
    for _ in range(12):
        size = [int(rng.integers(max(4, d // 12), max(6, d // 3))) for d in dims]
        size[1] = min(size[1], int(dims[1] * 0.6))
        lo = [int(rng.integers(3, max(4, d - sz - 3))) for d, sz in zip(dims, size)]
        lo[1] = 3  # furniture stands on the floor
        _box_shell(occ, lo, [l + sz for l, sz in zip(lo, size)], 2)
    n_struct = int(occ.sum())
    if n_struct < target_voxels:  # plenoxel haze: random floaters in free space
        need = target_voxels - n_struct
        free = np.flatnonzero(~occ.ravel())
        pick = rng.choice(free, size=min(need, free.size), replace=False)
        occ.ravel()[pick] = True
    vox = np.argwhere(occ).astype(np.int32)
    if vox.shape[0] > target_voxels:
        vox = vox[np.sort(rng.choice(vox.shape[0], size=target_voxels, replace=False))]
    return vox


def room_batch(seed: int, n_scenes: int, target_voxels: int, channels: int = 27, num_classes: int = 20,
               ignore_label: int = 255, shuffle: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(coords float32 [N,4], feats float32 [N,C], labels int64 [N]) for a batch of rooms."""
    rng = np.random.default_rng(seed)
    cs, fs, ls = [], [], []
    for b in range(n_scenes):
        vox = room_scene(rng, target_voxels)
        # The reference's loaders hand voxels over in the raster order of the plenoxel grid they were
        # decoded from (np.nonzero of `links`, co3d_3d/src/data/co3d.py:164-172, scannet.py:558-583;
        # crops / dropout are boolean masks and keep that order).  shuffle=True is the adversarial
        # case: neighbouring voxels land far apart in memory and gathers miss the L2.
        if shuffle:
            vox = vox[rng.permutation(vox.shape[0])]
        xyz = vox.astype(np.float32) + rng.random(vox.shape, dtype=np.float32) * np.float32(0.999)
        c = np.empty((vox.shape[0], 4), np.float32)
        c[:, 0] = b
        c[:, 1:] = xyz
        lab = rng.integers(0, num_classes, size=vox.shape[0]).astype(np.int64)
        lab[rng.random(vox.shape[0]) < 0.1] = ignore_label
        cs.append(c)
        fs.append(sh_features(rng, vox.shape[0], channels))
        ls.append(lab)
    return np.concatenate(cs), np.concatenate(fs), np.concatenate(ls)


def compact_records(coords: np.ndarray, feats: np.ndarray, labels: np.ndarray, ignore_label: int = 255):
    """The batch as PeRFception stores it on disk (co3d.py:126-172): per scene int32 `links` (flat cell index of a
    bounding grid, raster order), uint8 SH codes (`feats = u8 * 2/255 - 1` inverted exactly) and uint8 labels —
    32 bytes per voxel instead of the 132 of float coordinates + float features + int64 labels.  Decoding with
    `pipeline.plenoxel_decode(..., affine = identity + 0.5)` gives back coordinates inside the same voxels and
    bit-identical features.  Returns [(batch_index, links, sh_u8, labels_u8, reso), ...]."""
    out = []
    b_col = coords[:, 0].astype(np.int64)
    for b in np.unique(b_col):
        sel = b_col == b
        ijk = np.floor(coords[sel, 1:]).astype(np.int64)
        if ijk.min() < 0:
            raise ValueError("compact_records expects non-negative lattice coordinates")
        reso = tuple(int(v) + 1 for v in ijk.max(0))
        links = (ijk[:, 0] * reso[1] + ijk[:, 1]) * reso[2] + ijk[:, 2]
        if links.max() >= 2 ** 31:
            raise ValueError("grid too large for int32 links")
        u8 = np.rint((feats[sel].astype(np.float64) + 1.0) * 255.0 / 2.0)
        if u8.min() < 0 or u8.max() > 255:
            raise ValueError("features are not u8 * 2/255 - 1 codes")
        lab = labels[sel]
        if ((lab < 0) | (lab > 255)).any():
            raise ValueError("labels do not fit uint8")
        out.append((int(b), links.astype(np.int32), u8.astype(np.uint8), lab.astype(np.uint8), reso))
    return out


def faithful_room_batch(seed: int, n_scenes: int, target_voxels: int, scene_scale: float = 0.34, channels: int = 27,
                        num_classes: int = 20, ignore_label: int = 255, reso: int = 256, voxel_size: float = 0.02,
                        downsample_stride: int = 2) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """SURVEY.md §8d "config 2B": the coordinates the reference's PlenoxelScannetDataset really produces.

    A room is drawn on a reso^3 plenoxel grid, thinned to the lattice `c % downsample_stride == 0`
    (scannet.py:546-549, mode 1), and mapped like `load_data` does (scannet.py:612-615):
        xyz = (c / reso * 2 - 1) / scene_scale / voxel_size          (float32, torch semantics)
    so neighbouring samples sit `2 * downsample_stride / (reso * scene_scale * voxel_size)` voxels apart (2.3 at the
    median scene_scale 0.34): at tensor stride 1 the 3^3 kernel map is (almost) the centre tap only, real
    neighbourhoods appear at strides 2-4.  Rooms are tiled side by side along x until the scene holds
    `target_voxels` points.  Returns (coords float32 [N,4], feats float32 [N,C], labels int64 [N])."""
    rng = np.random.default_rng(seed)
    cs, fs, ls = [], [], []
    pitch = 2.0 * downsample_stride / (reso * scene_scale * voxel_size)
    for b in range(n_scenes):
        parts, have, tile = [], 0, 0
        while have < target_voxels:
            # one plenoxel grid: a room of reso x (0.43 reso) x reso cells, centred in the cube
            dims = (reso, max(12, int(round(reso * 130 / 300))), reso)
            cells_wanted = int(dims[0] * dims[1] * dims[2] * 0.2)        # ~1.5 M occupied cells per grid
            vox = room_scene(rng, cells_wanted, base=tuple(d * np.sqrt(1.0e6 / cells_wanted) for d in dims))
            vox = vox + np.array([0, (reso - dims[1]) // 2, 0], np.int32)
            vox = vox[(vox % downsample_stride == 0).all(1)]
            xyz = ((vox.astype(np.float32) / np.float32(reso) * np.float32(2) - np.float32(1))
                   / np.float32(scene_scale) / np.float32(voxel_size)).astype(np.float32)
            xyz[:, 0] += np.float32(tile * (reso // downsample_stride + 4) * pitch)     # next room beside the last
            parts.append(xyz)
            have += xyz.shape[0]
            tile += 1
        xyz = np.concatenate(parts)[:target_voxels]
        c = np.empty((xyz.shape[0], 4), np.float32)
        c[:, 0] = b
        c[:, 1:] = xyz
        lab = rng.integers(0, num_classes, size=xyz.shape[0]).astype(np.int64)
        lab[rng.random(xyz.shape[0]) < 0.1] = ignore_label
        cs.append(c)
        fs.append(sh_features(rng, xyz.shape[0], channels))
        ls.append(lab)
    return np.concatenate(cs), np.concatenate(fs), np.concatenate(ls)


def co3d_object(rng: np.random.Generator, lattice: int = 128) -> np.ndarray:
    """Integer voxels [n,3]: ellipsoid shell (thickness 3) + 3 solid blobs on a lattice^3 grid."""
    g = np.arange(lattice, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    ctr = lattice / 2.0
    scale = lattice / 128.0
    ax = rng.uniform(20, 55, size=3) * scale
    r = np.sqrt(((X - ctr) / ax[0]) ** 2 + ((Y - ctr) / ax[1]) ** 2 + ((Z - ctr) / ax[2]) ** 2)
    thick = 3.0 / float(ax.mean())
    occ = np.abs(r - 1.0) < thick / 2
    for _ in range(3):
        c = rng.uniform(ctr - 30 * scale, ctr + 30 * scale, size=3)
        rad = rng.uniform(4, 10) * scale
        occ |= ((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) < rad * rad
    return np.argwhere(occ).astype(np.int32)


def co3d_batch(seed: int, n_objects: int, channels: int = 27, num_classes: int = 51, lattice: int = 128,
               max_points: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(coords float32 [N,4] with float xyz after the emulated augmentation, feats [N,C], labels [B])."""
    rng = np.random.default_rng(seed)
    cs, fs = [], []
    for b in range(n_objects):
        vox = co3d_object(rng, lattice).astype(np.float32)
        if max_points and vox.shape[0] > max_points:
            vox = vox[np.sort(rng.choice(vox.shape[0], size=max_points, replace=False))]
        ctr = lattice / 2.0
        th = rng.uniform(0, 2 * np.pi)
        rot = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
        sc = np.float32(rng.uniform(0.6, 1.4))
        tr = rng.uniform(-0.2, 0.2, size=3).astype(np.float32) * np.float32(lattice / 2.0)
        xyz = ((vox - ctr) @ rot.T) * sc + ctr + tr
        c = np.empty((vox.shape[0], 4), np.float32)
        c[:, 0] = b
        c[:, 1:] = xyz
        cs.append(c)
        fs.append(sh_features(rng, vox.shape[0], channels))
    labels = rng.integers(0, num_classes, size=n_objects).astype(np.int64)
    return np.concatenate(cs), np.concatenate(fs), labels


def random_cloud(seed: int, n: int, extent: int = 12, n_batch: int = 2, channels: int = 8,
                 negative: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Small dense-ish float cloud with duplicates and negative coordinates (parity tests)."""
    rng = np.random.default_rng(seed)
    lo = -extent if negative else 0
    c = np.empty((n, 4), np.float32)
    c[:, 0] = rng.integers(0, n_batch, size=n)
    c[:, 1:] = rng.uniform(lo, extent, size=(n, 3)).astype(np.float32)
    f = rng.standard_normal((n, channels)).astype(np.float32)
    return c, f
