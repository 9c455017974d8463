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
                    batch_index: int = 0, affine: Optional[Sequence[float]] = None):
    """links [n] int32/int64 (flat cell indices), sh_u8 [n, C] uint8 -> coords [n,4] float32 (b,i,j,k), feats [n,C]
    float32 = sh * scale + min.  `affine` = 12 floats (row-major 3x3, then translation) applied to (i,j,k)."""
    lib = L.load()
    if links.dtype not in (torch.int32, torch.int64) or links.dim() != 1:
        raise RuntimeError("links must be a 1-D int32 / int64 tensor")
    if sh_u8.dtype != torch.uint8 or sh_u8.dim() != 2 or sh_u8.shape[0] != links.shape[0]:
        raise RuntimeError("sh must be uint8 [n, C]")
    links, sh_u8 = links.contiguous(), sh_u8.contiguous()
    n, C = sh_u8.shape
    dev = links.device
    coords = torch.empty((n, 4), dtype=torch.float32, device=dev)
    feats = torch.empty((n, C), dtype=torch.float32, device=dev)
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

    def compute(self):
        seen, correct, positive = (self.counts[i].to(torch.float32) for i in range(3))
        present = seen != 0
        ious = torch.where(present, correct / (seen + positive - correct).clamp_min(1), torch.zeros_like(seen))
        accs = torch.where(present, correct / seen.clamp_min(1), torch.zeros_like(seen))
        if self.void_label is not None:
            return ious[:-1].mean(), ious, accs[:-1].mean(), accs
        return ious.mean(), ious, accs.mean(), accs
