"""Functional CPU restatement of the two benchmark networks on top of oracle/ref_ops.py.

TEST INFRASTRUCTURE ONLY (see ref_ops.py header; parity unpinned against ME itself).
Follows the forward passes of the reference models:
  ResNetBase.forward      co3d_3d/src/models/mink/resnet.py:163-177
  BasicBlockBase.forward  co3d_3d/src/models/mink/modules/resnet_block.py:53-69
  Res16UNet.forward       co3d_3d/src/models/mink/res16unet.py:391-435
Parameters come from a state_dict with the reference's key names (conv `kernel`/`bias`,
`bn.weight` ... ), given as torch CPU tensors; autograd through these functions provides the
gradient oracle.  BatchNorm always uses batch statistics (training mode).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import ref_ops as R

ONE = (1, 1, 1)
TWO = (2, 2, 2)


class _Net:
    def __init__(self, params: Dict[str, torch.Tensor], mgr: R.OracleManager):
        self.p = params
        self.m = mgr

    def conv(self, name, x, ts, k, stride=1):
        w = self.p[name + ".kernel"]
        b = self.p.get(name + ".bias")
        if w.dim() == 2:  # kernel_volume 1, stride 1 -> plain matrix product (sparse_conv.py:391-395)
            out = x @ w
            return (out + b.view(1, -1) if b is not None else out), ts
        ts_out = self.m.stride(ts, (stride,) * 3)
        nbr = self.m.kernel_map(ts, ts_out, (k,) * 3)
        return R.conv_forward(x, w, nbr, b), ts_out

    def conv_tr(self, name, x, ts, k, stride):
        w = self.p[name + ".kernel"]
        ts_out = tuple(t // stride for t in ts)
        nbr = self.m.kernel_map(ts, ts_out, (k,) * 3, transpose=True)
        return R.conv_forward(x, w, nbr, None), ts_out

    def bn(self, name, x):
        return R.batch_norm(x, self.p[name + ".bn.weight"], self.p[name + ".bn.bias"], training=True)

    def block(self, name, x, ts, stride=1):
        out, ts_out = self.conv(name + ".conv1", x, ts, 3, stride)
        out = torch.relu(self.bn(name + ".norm1", out))
        out, _ = self.conv(name + ".conv2", out, ts_out, 3, 1)
        out = self.bn(name + ".norm2", out)
        if (name + ".downsample.0.kernel") in self.p:
            res, _ = self.conv(name + ".downsample.0", x, ts, 1, stride)
            res = self.bn(name + ".downsample.1", res)
        else:
            res = x
        return torch.relu(out + res), ts_out

    def stage(self, name, x, ts, stride=1):
        i = 0
        while (f"{name}.{i}.conv1.kernel") in self.p:
            x, ts = self.block(f"{name}.{i}", x, ts, stride if i == 0 else 1)
            i += 1
        return x, ts


def _input(coords_f: np.ndarray, feats: torch.Tensor, use_c: bool) -> Tuple[R.OracleManager, torch.Tensor]:
    mgr = R.OracleManager(coords_f, use_c=use_c)
    m = mgr.maps[ONE].shape[0]
    return mgr, R.segment_mean(feats, mgr.inverse, m, "mean")


def resnet_forward(params, coords_f: np.ndarray, feats: torch.Tensor, use_c: bool = False):
    """ResNet14/18/34 logits [B, out]; rows ordered by batch index."""
    mgr, x = _input(coords_f, feats, use_c)
    n = _Net(params, mgr)
    x, ts = n.conv("conv1", x, ONE, 3)
    x = torch.relu(n.bn("bn1", x))
    ts2 = mgr.stride(ts, TWO)
    x = R.sum_pool(x, mgr.kernel_map(ts, ts2, TWO))
    ts = ts2
    for i in range(1, 5):
        x, ts = n.stage(f"layer{i}", x, ts, stride=2)
    n_batch = int(mgr.maps[ONE][:, 0].max()) + 1
    x = R.global_avg_pool(x, mgr.maps[ts], n_batch)
    out, _ = n.conv("final", x, ts, 1)
    return out


def resunet_forward(params, coords_f: np.ndarray, feats: torch.Tensor, use_c: bool = False):
    """Res16UNet per-point logits [N, out] (sliced back with the inverse map)."""
    mgr, x = _input(coords_f, feats, use_c)
    n = _Net(params, mgr)
    ts = ONE
    x, _ = n.conv("conv0p1s1.0", x, ts, 3)
    x = torch.relu(n.bn("conv0p1s1.1", x))
    x, _ = n.conv("conv0p1s1.3", x, ts, 3)
    x = torch.relu(n.bn("conv0p1s1.4", x))
    skips = [(x, ts)]
    for i, name in enumerate(["conv1p1s2", "conv2p2s2", "conv3p4s2", "conv4p8s2"]):
        x, ts = n.conv(name + ".0", x, ts, 2, 2)
        x = torch.relu(n.bn(name + ".1", x))
        x, ts = n.stage(f"block{i + 1}", x, ts)
        skips.append((x, ts))
    skips.pop()
    for i, name in enumerate(["convtr4p16s2", "convtr5p8s2", "convtr6p4s2", "convtr7p2s2"]):
        x, ts = n.conv_tr(name + ".0", x, ts, 2, 2)
        x = torch.relu(n.bn(name + ".1", x))
        s, s_ts = skips.pop()
        assert s_ts == ts
        x = torch.cat([x, s], dim=1)
        x, ts = n.stage(f"block{5 + i}", x, ts)
    x, _ = n.conv("final", x, ts, 1)
    return x.index_select(0, torch.from_numpy(mgr.inverse.astype(np.int64)))
