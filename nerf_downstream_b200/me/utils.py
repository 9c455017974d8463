"""MinkowskiEngine.utils equivalents used by the reference's data path
(co3d_3d/src/data/utils.py:25-50 -> sparse_collate(..., dtype=torch.float32))."""
from collections.abc import Sequence

import numpy as np
import torch


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """Concatenate per-sample [n_i, D] coordinates into [N, D+1] with the batch index in column 0.
    int32 -> coordinates are floored first; float32 keeps them as they are (appendix A.2)."""
    assert isinstance(coords, Sequence) and len(coords) > 0, "coords must be a non-empty sequence"
    assert dtype in (torch.int32, torch.float32), "Only torch.int32, torch.float32 supported for coordinates."
    D = int(np.asarray(coords[0]).shape[1]) if not isinstance(coords[0], torch.Tensor) else coords[0].shape[1]
    N = sum(int(c.shape[0]) for c in coords)
    out_dev = device
    if out_dev is None and isinstance(coords[0], torch.Tensor):
        out_dev = coords[0].device
    bcoords = torch.zeros((N, D + 1), dtype=dtype, device=out_dev)
    s = 0
    for b, c in enumerate(coords):
        if not isinstance(c, torch.Tensor):
            c = torch.from_numpy(np.asarray(c))
        assert c.shape[1] == D, "all samples must have the same coordinate dimension"
        if dtype == torch.int32:
            c = torch.floor(c.double()).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
        else:
            c = c.to(torch.float32)
        n = c.shape[0]
        bcoords[s:s + n, 1:] = c.to(bcoords.device)
        bcoords[s:s + n, 0] = b
        s += n
    return bcoords


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """(batched coordinates, concatenated features[, concatenated labels])."""
    use_label = labels is not None
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)

    def _cat(items):
        ts = [torch.from_numpy(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in items]
        out = torch.cat(ts, 0)
        return out.to(device) if device is not None else out

    feats_batch = _cat(feats)
    if use_label:
        return bcoords, feats_batch, _cat(labels)
    return bcoords, feats_batch


def batch_sparse_collate(data, dtype=torch.int32, device=None):
    return sparse_collate(*list(zip(*data)), dtype=dtype, device=device)


class SparseCollation:
    def __init__(self, limit_numpoints=-1, dtype=torch.int32, device=None):
        self.limit_numpoints = limit_numpoints
        self.dtype = dtype
        self.device = device

    def __call__(self, list_data):
        coords, feats, labels = list(zip(*list_data))
        if self.limit_numpoints > 0:
            keep, total = [], 0
            for i, c in enumerate(coords):
                total += len(c)
                if total > self.limit_numpoints and i > 0:
                    break
                keep.append(i)
            coords = [coords[i] for i in keep]
            feats = [feats[i] for i in keep]
            labels = [labels[i] for i in keep]
        return sparse_collate(coords, feats, labels, dtype=self.dtype, device=self.device)


def sparse_quantize(*args, **kwargs):
    raise NotImplementedError(
        "ME.utils.sparse_quantize is used only by out-of-scope datasets (SURVEY.md §2.1 row 11); "
        "use TensorField(...).sparse() for on-device quantisation")


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME's kaiming init for kernels shaped (K, Cin, Cout)."""
    import math
    if tensor.dim() == 3:
        k, cin, cout = tensor.shape
        fan = k * cin if mode == "fan_in" else k * cout
    else:
        cin, cout = tensor.shape
        fan = cin if mode == "fan_in" else cout
    gain = torch.nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
