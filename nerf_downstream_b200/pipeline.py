"""The two ends of the hot path on the GPU (SURVEY.md §8f "next" rows 2 and 3).

`plenoxel_decode` turns one PeRFception plenoxel record (what `COD3D.load_data` reads from disk,
co3d_3d/src/data/co3d.py:126-172) into the `(coordinates, features)` pair the reference builds on the CPU in
`__getitem__` (co3d.py:196-203) and hands to `ME.TensorField`; `IoUMeter` is the reference's metric
(co3d_3d/src/metrics.py:5-58) with the per-class Python loop replaced by one kernel over the logits.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import lib as L


def plenoxel_decode(links: torch.Tensor, sh_u8: torch.Tensor, sh_scale: float, sh_min: float, reso: Sequence[int],
                    batch_index: int = 0, affine: Optional[Sequence[float]] = None,
                    out_coords: Optional[torch.Tensor] = None, out_feats: Optional[torch.Tensor] = None,
                    rows: Optional[torch.Tensor] = None):
    """links [n] int32/int64 (flat cell indices), sh_u8 [n, C] uint8 -> coords [n,4] float32 (b,i,j,k), feats [n,C]
    float32 = sh * scale + min.  `affine` = 12 floats (row-major 3x3, then translation) applied to (i,j,k).
    `out_coords` / `out_feats`: contiguous [n,4] / [n,C] float32 destinations (e.g. one scene's rows of a batch
    buffer: the records of a batch decode straight into the collated tensors, no torch.cat).
    `rows` (int32 [m]): decode only these records, output row j <- record rows[j] (`plenoxel_select_rows`: the
    reference's RandomCrop / CoordinateDropout as a row list instead of boolean-mask copies of decoded tensors)."""
    if rows is not None:
        return _plenoxel_decode_rows(links, sh_u8, sh_scale, sh_min, reso, batch_index, affine, rows)
    lib = L.load()
    if links.dtype not in (torch.int32, torch.int64) or links.dim() != 1:
        raise RuntimeError("links must be a 1-D int32 / int64 tensor")
    if sh_u8.dtype != torch.uint8 or sh_u8.dim() != 2 or sh_u8.shape[0] != links.shape[0]:
        raise RuntimeError("sh must be uint8 [n, C]")
    links, sh_u8 = links.contiguous(), sh_u8.contiguous()
    n, C = sh_u8.shape
    dev = links.device
    coords = out_coords if out_coords is not None else torch.empty((n, 4), dtype=torch.float32, device=dev)
    feats = out_feats if out_feats is not None else torch.empty((n, C), dtype=torch.float32, device=dev)
    if coords.shape != (n, 4) or feats.shape != (n, C) or coords.dtype != torch.float32 or feats.dtype != torch.float32:
        raise RuntimeError("decode destinations must be float32 [n,4] / [n,C]")
    r = (ctypes.c_int32 * 3)(*[int(v) for v in reso])
    aff = None
    if affine is not None:
        if len(affine) != 12:
            raise RuntimeError("affine must hold 12 floats (3x3 row-major + translation)")
        aff = (ctypes.c_float * 12)(*[float(v) for v in affine])
    L.check(lib.spc_plenoxel_decode(L.ptr(links), int(links.dtype == torch.int64), n, ctypes.cast(r, ctypes.c_void_p),
                                    int(batch_index), ctypes.cast(aff, ctypes.c_void_p) if aff is not None else None,
                                    L.ptr(sh_u8), C, float(sh_scale), float(sh_min), L.ptr(coords), L.ptr(feats),
                                    L.stream()), "spc_plenoxel_decode")
    return coords, feats


def _affine12(affine):
    if affine is None:
        return None
    if len(affine) != 12:
        raise RuntimeError("affine must hold 12 floats (3x3 row-major + translation)")
    return (ctypes.c_float * 12)(*[float(v) for v in affine])


def _plenoxel_decode_rows(links, sh_u8, sh_scale, sh_min, reso, batch_index, affine, rows):
    lib = L.load()
    if rows.dtype != torch.int32 or rows.dim() != 1:
        raise RuntimeError("rows must be a 1-D int32 tensor")
    links, sh_u8, rows = links.contiguous(), sh_u8.contiguous(), rows.contiguous()
    m, C = rows.shape[0], sh_u8.shape[1]
    coords = torch.empty((m, 4), dtype=torch.float32, device=links.device)
    feats = torch.empty((m, C), dtype=torch.float32, device=links.device)
    r = (ctypes.c_int32 * 3)(*[int(v) for v in reso])
    aff = _affine12(affine)
    L.check(lib.spc_plenoxel_decode_rows(L.ptr(links), int(links.dtype == torch.int64), L.ptr(rows), m,
                                         ctypes.cast(r, ctypes.c_void_p), int(batch_index),
                                         ctypes.cast(aff, ctypes.c_void_p) if aff is not None else None, L.ptr(sh_u8), C,
                                         float(sh_scale), float(sh_min), L.ptr(coords), L.ptr(feats), L.stream()),
            "spc_plenoxel_decode_rows")
    return coords, feats


def plenoxel_crop_rows(links: torch.Tensor, reso: Sequence[int], size3: Sequence[float], u3: Sequence[float],
                       rows: Optional[torch.Tensor] = None, affine: Optional[Sequence[float]] = None):
    """One draw of RandomCrop (transforms.py:206-226) over the records `rows` (None: all) of a plenoxel record, the
    lattice coordinates mapped through `affine` first.  Returns (kept record numbers int32 [k], fits): `fits` = the box
    covers the extent on every axis (the reference returns its input unchanged).  One host synchronisation."""
    lib = L.load()
    links = links.contiguous()
    n = int(rows.shape[0]) if rows is not None else int(links.shape[0])
    out = torch.empty(max(n, 1), dtype=torch.int32, device=links.device)
    res = torch.empty(2, dtype=torch.int32, device=links.device)
    ws_bytes = int(lib.spc_plenoxel_crop_workspace(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=links.device)
    r = (ctypes.c_int32 * 3)(*[int(v) for v in reso])
    aff = _affine12(affine)
    u = (ctypes.c_float * 3)(*[float(v) for v in u3])
    sz = (ctypes.c_float * 3)(*[float(v) for v in size3])
    L.check(lib.spc_plenoxel_crop_select(L.ptr(links), int(links.dtype == torch.int64),
                                         L.ptr(rows.contiguous()) if rows is not None else None, n,
                                         ctypes.cast(r, ctypes.c_void_p),
                                         ctypes.cast(aff, ctypes.c_void_p) if aff is not None else None,
                                         ctypes.cast(u, ctypes.c_void_p), ctypes.cast(sz, ctypes.c_void_p), L.ptr(out),
                                         L.ptr(res), L.ptr(ws), ws_bytes, L.stream()), "spc_plenoxel_crop_select")
    kept, fits = res.tolist()
    return out[:kept], bool(fits)


def plenoxel_select_rows(links: torch.Tensor, reso: Sequence[int], steps: Sequence[tuple],
                         affine: Optional[Sequence[float]] = None) -> Optional[torch.Tensor]:
    """The point-dropping transformations of the reference's train lists as ONE row list of the record, drawn in the
    reference's RNG order (python `random` for the application ratios, numpy for the draws):
      ("RandomCrop", dict(x=, y=, z=, application_ratio=1, max_retries=10))        transforms.py:195-243
      ("CoordinateDropout", dict(dropout_ratio=0.2, application_ratio=0.2))        transforms.py:247-264
    `affine`: the affine chain in front of these steps.  Returns int32 record numbers in the reference's row order, or
    None when every step left the rows as they were.  The record is then decoded once for those rows
    (`plenoxel_decode(..., rows=...)`), instead of being decoded in full and copied through boolean masks."""
    import random as py_random

    import numpy as np
    rows = None
    n = int(links.shape[0])
    for name, kw in steps:
        cur = n if rows is None else int(rows.shape[0])
        if name == "RandomCrop":
            if py_random.random() > kw.get("application_ratio", 1):
                continue
            size3 = (kw["x"], kw["y"], kw["z"])
            retries, chosen = 0, None
            while True:
                state = np.random.get_state()
                u3 = np.random.rand(1, 3)[0]
                kept, fits = plenoxel_crop_rows(links, reso, size3, u3, rows, affine)
                if fits:                      # `np.prod(coord_range == 0)`: the reference returns BEFORE its first draw
                    np.random.set_state(state)
                    break
                if kept.shape[0] > 0:
                    chosen = kept
                    break
                retries += 1
                if retries >= kw.get("max_retries", 10):
                    break
            if chosen is not None:
                rows = chosen
        elif name == "CoordinateDropout":
            if py_random.random() < kw.get("application_ratio", 0.2):
                inds = np.random.choice(cur, int(cur * (1 - kw.get("dropout_ratio", 0.2))), replace=False)
                idx = torch.from_numpy(inds.astype(np.int64)).to(links.device)
                rows = idx.to(torch.int32) if rows is None else rows[idx]
        else:
            raise KeyError(f"{name} does not drop points (RandomCrop / CoordinateDropout do)")
    return rows


def plenoxel_decode_augmented(links: torch.Tensor, sh_u8: torch.Tensor, sh_scale: float, sh_min: float,
                              reso: Sequence[int], transformations: Sequence[str], batch_index: int = 0,
                              params: Optional[dict] = None):
    """Decode a plenoxel record AND apply the reference's affine-type train transformations (RandomRotation,
    RandomAffine, RandomHorizontalFlip, RandomTranslation, CoordinateUniformTranslation, RandomScale,
    DimensionlessCoordinates; co3d_3d/src/data/transforms.py:284-460) in the same kernel pass: the chain is drawn on
    the host in the reference's RNG order (`augment.sample_chain`), composed into one 3x3 + translation and handed to
    `spc_plenoxel_decode` as `affine12` — instead of one numpy pass per transform on the CPU.  A horizontal flip
    mirrors about the data's maximum, which costs one extra coordinates-only decode + a reduction.
    Returns (coords [n,4] float32, feats [n,C] float32, the AffineChain)."""
    from . import augment
    raw = []

    def axis_max(chain, axis):
        if not raw:
            raw.append(plenoxel_decode(links, sh_u8[:, :0], 1.0, 0.0, reso, batch_index)[0][:, 1:].double())
        a = torch.tensor(chain.A[axis], dtype=torch.float64, device=raw[0].device)
        return float((raw[0] @ a).max().item() + chain.t[axis])

    chain = augment.sample_chain(transformations, axis_max=axis_max, params=params)
    coords, feats = plenoxel_decode(links, sh_u8, sh_scale, sh_min, reso, batch_index,
                                    chain.as_affine12() if chain.steps else None)
    return coords, feats, chain


# ---- record -> sample: the reference's Dataset.__getitem__ after the file has been read ------------------------------
SCANNET_VALID_CLASS_IDS = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39)   # scannet.py:43-64
SCANNET_NUM_LABELS = 41                                                                              # scannet.py:451


def scannet_label_map(ignore_label: int = -100, void_label: Optional[int] = None) -> dict:
    """NYU40 id -> train id (0..19), everything else -> ignore_label, optional void class = 20 (scannet.py:518-528)."""
    label_map, n_used = {}, 0
    for lab in range(SCANNET_NUM_LABELS):
        if lab in SCANNET_VALID_CLASS_IDS:
            label_map[lab] = n_used
            n_used += 1
        else:
            label_map[lab] = ignore_label
    label_map[ignore_label] = ignore_label
    if void_label is not None and void_label != ignore_label:
        label_map[void_label] = n_used
    return label_map


def _select_features(features: Sequence[str], available: dict) -> torch.Tensor:
    """`features.append(eval(f))` over the names of configs/feature_*.gin (co3d.py:217-220, scannet.py:633-636)."""
    missing = [f for f in features if f not in available]
    if missing:
        raise KeyError(f"unknown feature name(s) {missing}; this sample offers {sorted(available)}")
    return torch.cat([available[f] for f in features], dim=1).float()


def co3d_sample(links: torch.Tensor, density: torch.Tensor, sh_u8: torch.Tensor, sh_scale: float, sh_min: float,
                reso: Sequence[int], features: Sequence[str] = ("sh",), transformations: Sequence[str] = (),
                params: Optional[dict] = None) -> dict:
    """`Co3DDatasetBase.__getitem__` (co3d.py:178-235) for one record already on the GPU: lattice coordinates and
    dequantised SH from `spc_plenoxel_decode`, the (per-point, as the reference writes it: `mean(dim=1)`) centred and
    max-norm-scaled `xyzs`, the train transformations, the configured feature columns."""
    from . import augment
    coords4, sh = plenoxel_decode(links, sh_u8, sh_scale, sh_min, reso)
    coordinates = coords4[:, 1:]
    xyzs = coordinates - coordinates.mean(dim=1, keepdim=True)
    xyzs = xyzs / torch.linalg.norm(xyzs, dim=1).max()
    raw = torch.cat([xyzs, density.to(sh.device).float().view(-1, 1), sh], dim=1).float()
    if len(transformations) > 0:
        coordinates, raw, _ = augment.apply_transformations(transformations, coordinates, raw, None, params)
    xyzs, dens, shf = raw[:, :3], raw[:, 3:4], raw[:, 4:]
    feats = _select_features(features, {"xyzs": xyzs, "density": dens, "sh": shf, "ones": torch.ones_like(dens)})
    return {"coordinates": coordinates, "features": feats, "xyzs": xyzs}


def scannet_plenoxel_sample(links: torch.Tensor, density: torch.Tensor, sh_u8: torch.Tensor, sh_scale: float,
                            sh_min: float, reso: Sequence[int], labels: torch.Tensor, dists: torch.Tensor,
                            scene_scale: float, voxel_size: float = 0.02, downsample_stride: int = 2,
                            ignore_label: int = -100, void_label: Optional[int] = None, valid_thres: float = 0.05,
                            ignore_thres: Optional[float] = None, features: Sequence[str] = ("sh",),
                            transformations: Sequence[str] = (), params: Optional[dict] = None) -> dict:
    """`PlenoxelScannetDataset.load_data` + `__getitem__` (scannet.py:558-654) for one record on the GPU:
    void labelling by surface distance, optional distance filter, decode, thinning to the `c % stride == 0` lattice
    (downsample mode 1), `xyz = (c / reso * 2 - 1) / scene_scale / voxel_size`, train transformations, feature
    columns, NYU40 -> 20-class label map."""
    from . import augment
    dev = links.device
    labels = labels.to(dev).long().clone().view(-1)
    dists = dists.to(dev).float().view(-1)
    density = density.to(dev).float().view(-1, 1)
    labels[dists > valid_thres] = void_label if void_label is not None else ignore_label
    if ignore_thres is not None and ignore_thres > 0:
        keep = dists < ignore_thres
        links, sh_u8, density, labels = links[keep], sh_u8[keep], density[keep], labels[keep]
        # (the reference filters `dists` only implicitly — scannet.py:574-579 leaves it unfiltered and would fail in
        #  `torch.cat`; here the distances follow their voxels)
        dists = dists[keep]
    if len(features) > 1:                                     # scannet.py:593-594
        density = density / (density.abs().max() + 1e-5)
    coords4, sh = plenoxel_decode(links, sh_u8, sh_scale, sh_min, reso)
    coordinates = coords4[:, 1:]
    sel = (coordinates % downsample_stride == 0).all(dim=1)
    coordinates, sh, density, labels, dists = coordinates[sel], sh[sel], density[sel], labels[sel], dists[sel]
    r = torch.tensor([float(v) for v in reso], dtype=torch.float32, device=dev)
    xyzs = (coordinates / r * 2 - 1.0) / scene_scale / voxel_size
    raw = torch.cat([xyzs, dists.view(-1, 1), density, sh], dim=1).float()
    if len(transformations) > 0:
        xyzs, raw, labels = augment.apply_transformations(transformations, xyzs, raw, labels, params)
    dist_f, dens, shf = raw[:, 3:4], raw[:, 4:5], raw[:, 5:]
    feats = _select_features(features, {"dists": dist_f, "density": dens, "sh": shf, "ones": torch.ones_like(dens),
                                        "xyzs": raw[:, :3]})
    mapped = torch.full_like(labels, ignore_label)            # `label_map[x]` (scannet.py:637-640)
    for k, v in scannet_label_map(ignore_label, void_label).items():
        mapped = torch.where(labels == k, torch.full_like(labels, v), mapped)
    return {"coordinates": xyzs.float(), "features": feats, "xyzs": xyzs.float(), "labels": mapped.to(torch.int32),
            "dists": dist_f}


def seg_counts(logits: torch.Tensor, target: torch.Tensor, ignore_label: int, out: Optional[torch.Tensor] = None):
    """counts[3, C] int64 (+= when `out` is given): per class #seen, #correct, #predicted of argmax(logits)."""
    lib = L.load()
    if logits.dtype != torch.float32 or logits.dim() != 2:
        raise RuntimeError("logits must be float32 [n, C]")
    if target.dtype != torch.int64 or target.shape != (logits.shape[0],):
        raise RuntimeError("target must be int64 [n]")
    logits, target = logits.contiguous(), target.contiguous()
    n, C = logits.shape
    if out is None:
        out = torch.zeros((3, C), dtype=torch.int64, device=logits.device)
    L.check(lib.spc_seg_metrics(L.ptr(logits), L.ptr(target), n, C, int(ignore_label), L.ptr(out), L.stream()),
            "spc_seg_metrics")
    return out


def seg_head_loss(out, field, labels: torch.Tensor, ignore_index: int = -100, weight: Optional[torch.Tensor] = None,
                  counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`SegLoss(out.slice(field).F, labels)` (res16unet.py:435 + segmentation_training.py:27-44) without materialising
    the sliced logits: one kernel gathers each point's voxel row through the field's inverse map, evaluates the
    (class-weighted, ignore-aware) cross-entropy, accumulates the gradient on the voxel rows and — when `counts`
    [3, C] int64 is given — the IoUMeter counts of the same argmax.  `out` is the network's SparseTensor head output
    (`models.SparseResUNet.forward_sparse`), `field` the TensorField it was computed from."""
    from . import ops
    inv = field.inverse_mapping(out.coordinate_map_key)
    return ops.seg_head(out.F, inv, labels, ignore_index, weight, counts)


class IoUMeter:
    """co3d_3d/src/metrics.py IoUMeter: update() accumulates, compute() -> (miou, ious, mAcc, accs)."""

    def __init__(self, num_classes: int, ignore_label: int, void_label=None):
        self.num_classes, self.ignore_label, self.void_label = num_classes, ignore_label, void_label
        self.counts = None

    def update(self, logits: torch.Tensor, targets: torch.Tensor):
        if self.counts is None:
            self.counts = torch.zeros((3, self.num_classes), dtype=torch.int64, device=logits.device)
        seg_counts(logits, targets, self.ignore_label, out=self.counts)

    def counts_buffer(self, device) -> torch.Tensor:
        """The [3, C] int64 accumulator, for kernels that add to it directly (`seg_head_loss(counts=...)`)."""
        if self.counts is None:
            self.counts = torch.zeros((3, self.num_classes), dtype=torch.int64, device=device)
        return self.counts

    def all_reduce(self, group=None) -> None:
        """Sum the counts over the ranks of a data-parallel job (`dist_reduce_fx="sum"`, metrics.py:17-28)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.counts is None:
                # a rank with an empty validation shard (or whose every batch failed) must still enter the collective
                dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
                self.counts = torch.zeros((3, self.num_classes), dtype=torch.int64, device=dev)
            dist.all_reduce(self.counts, op=dist.ReduceOp.SUM, group=group)

    def compute(self):
        seen, correct, positive = (self.counts[i].to(torch.float32) for i in range(3))
        present = seen != 0
        ious = torch.where(present, correct / (seen + positive - correct).clamp_min(1), torch.zeros_like(seen))
        accs = torch.where(present, correct / seen.clamp_min(1), torch.zeros_like(seen))
        if self.void_label is not None:
            return ious[:-1].mean(), ious, accs[:-1].mean(), accs
        return ious.mean(), ious, accs.mean(), accs
