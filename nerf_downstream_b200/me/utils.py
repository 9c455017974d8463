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


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cuda"):
    """ME.utils.sparse_quantize as the reference calls it (co3d_3d/src/data/scannet.py:235-242): voxelise a point
    cloud — `floor(coordinates / quantization_size)`, one row per occupied voxel.

    Returns, like ME, `discrete_coords[, features][, labels][, unique_index][, inverse_index]` (numpy arrays for numpy
    input, tensors on the GPU otherwise).  A voxel whose points carry different labels gets `ignore_label`.  Voxels
    come in the order of their first point (ME's CPU hash map leaves the order unspecified); `unique_index` is that
    first point.  The hashing runs in the coordinate kernels (`spc_coords_insert`) — there is no CPU path, so this
    needs a CUDA device even though ME's default is "cpu"."""
    from .. import lib as L
    from .. import ops
    if torch.is_tensor(coordinates) and coordinates.is_cuda:
        dev = coordinates.device
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("sparse_quantize runs in the CUDA coordinate kernels (no CPU backend)")
        dev = torch.device("cuda", torch.cuda.current_device())
    return _sparse_quantize(coordinates, features, labels, ignore_label, return_index, return_inverse,
                            return_maps_only, quantization_size, dev,
                            lambda disc: ops.coords_insert(disc, L.SRC_INT, (1, 1, 1)))


def _sparse_quantize(coordinates, features, labels, ignore_label, return_index, return_inverse, return_maps_only,
                     quantization_size, dev, insert):
    """Host logic of `sparse_quantize`; `insert(int32 [N,4]) -> (CoordMap, first, inverse, count)` is the hash kernel."""
    import numpy as np
    is_np = isinstance(coordinates, np.ndarray)

    def to_dev(a):
        if a is None:
            return None
        return (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a).to(dev)

    c = to_dev(coordinates)
    assert c.dim() == 2, "The coordinates must be a 2D matrix. The shape of the input is " + str(tuple(c.shape))
    if features is not None:
        assert features.ndim == 2 and coordinates.shape[0] == features.shape[0]
    if labels is not None:
        assert coordinates.shape[0] == len(labels)
    n, D = c.shape
    if D != 3:
        raise NotImplementedError("this engine is built for 3 spatial dimensions")
    if quantization_size is not None:
        q = quantization_size
        if isinstance(q, (list, tuple, np.ndarray, torch.Tensor)):
            assert len(q) == D, "Quantization size and coordinates size mismatch."
            q = torch.as_tensor(np.asarray(q) if not torch.is_tensor(q) else q, device=dev).to(c.dtype if
                                                                                              c.is_floating_point() else torch.float64)
        elif isinstance(q, (int, float)):
            # a device tensor, not a Python scalar: torch divides by a host scalar as `c * (1 / q)`, numpy as `c / q`
            q = torch.tensor(q, dtype=c.dtype if c.is_floating_point() else torch.float64, device=dev)
        else:
            raise ValueError("Not supported type for quantization_size.")
        c = torch.floor(c / q)                      # true division in the input's precision, as numpy does
    elif c.is_floating_point():
        c = torch.floor(c)
    disc = torch.zeros((n, 4), dtype=torch.int32, device=dev)
    disc[:, 1:] = c.to(torch.int32)
    cmap, first, inverse, _count = insert(disc)
    unique_index, inverse_index = first.long(), inverse.long()

    def back(t):
        return t.cpu().numpy() if is_np else t

    if return_maps_only:
        out = [back(unique_index)]
        if return_inverse:
            out.append(back(inverse_index))
        return out[0] if len(out) == 1 else tuple(out)
    out = [back(cmap.coords[:, 1:].contiguous())]
    if features is not None:
        out.append(back(to_dev(features)[unique_index]))
    if labels is not None:
        lab = to_dev(labels).long()
        vox = lab[unique_index].clone()
        conflict = lab != vox[inverse_index]
        vox[inverse_index[conflict]] = ignore_label
        vox = vox.to(torch.int32)
        out.append(back(vox))
    if return_index:
        out.append(back(unique_index))
    if return_inverse:
        out.append(back(inverse_index))
    return out[0] if len(out) == 1 else tuple(out)


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
