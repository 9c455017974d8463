"""Op-level check INSIDE a real UNet backward pass: every conv dgrad/wgrad and BN backward is
recomputed in fp64 on the CPU from the exact tensors the CUDA op received."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nerf_downstream_b200 import me as ME  # noqa: E402
from nerf_downstream_b200 import models, ops, synth  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

voxels = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
torch.manual_seed(1)
ops.set_default_precision(mode)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300)), float((a - b).abs().max()), float(b.abs().max())


orig_dgrad, orig_wgrad, orig_fwd = ops.conv_dgrad_raw, ops.conv_wgrad_raw, ops.conv_fwd_raw
n = [0]


def chk_fwd(x, w, bias, km, precision):
    out = orig_fwd(x, w, bias, km, precision)
    nbr = km.nbr.cpu().numpy()
    ref = R.conv_forward(x.double().cpu(), w.double().cpu(), nbr, bias.double().cpu() if bias is not None else None)
    print("fwd   K=%2d %4d->%4d M=%6d rel=%.2e max=%.2e/%.2e" % (km.K, w.shape[1], w.shape[2], km.m_out, *rel(out, ref)))
    return out


def chk_dgrad(g, w, km, precision):
    out = orig_dgrad(g, w, km, precision)
    nbr = km.nbr.cpu().numpy()
    gd, wd = g.double().cpu(), w.double().cpu()
    ref = torch.zeros(km.m_in, w.shape[1], dtype=torch.float64)
    for k in range(km.K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if o.size:
            ref.index_add_(0, torch.from_numpy(nbr[k, o].astype(np.int64)), gd[torch.from_numpy(o)] @ wd[k].t())
    print("dgrad K=%2d %4d->%4d M=%6d rel=%.2e max=%.2e/%.2e" % (km.K, w.shape[1], w.shape[2], km.m_out, *rel(out, ref)))
    return out


def chk_wgrad(x, g, km, K, c_in, c_out, precision):
    out = orig_wgrad(x, g, km, K, c_in, c_out, precision)
    nbr = km.nbr.cpu().numpy()
    xd, gd = x.double().cpu(), g.double().cpu()
    ref = torch.zeros(K, c_in, c_out, dtype=torch.float64)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if o.size:
            ref[k] = xd[torch.from_numpy(nbr[k, o].astype(np.int64))].t() @ gd[torch.from_numpy(o)]
    print("wgrad K=%2d %4d->%4d M=%6d rel=%.2e max=%.2e/%.2e" % (K, c_in, c_out, km.m_out, *rel(out, ref)))
    return out


ops.conv_dgrad_raw, ops.conv_wgrad_raw = chk_dgrad, chk_wgrad
if "--fwd" in sys.argv:
    ops.conv_fwd_raw = chk_fwd

orig_bn_bwd = ops.BatchNormFn.backward


def bn_bwd(ctx, dy):
    res = orig_bn_bwd(ctx, dy)
    x, y, mean, var, gamma = ctx.saved_tensors
    eps, relu, use_batch, has_res, affine = ctx.cfg
    xd, gd, dyd = x.double().cpu(), gamma.double().cpu(), dy.double().cpu()
    if relu:
        dyd = dyd * (y.double().cpu() > 0)
    m = xd.shape[0]
    mu = xd.mean(0)
    v = xd.var(0, unbiased=False)
    istd = 1.0 / torch.sqrt(v + eps)
    xhat = (xd - mu) * istd
    dbeta = dyd.sum(0)
    dgamma = (dyd * xhat).sum(0)
    dx = gd * istd * (dyd - dbeta / m - xhat * dgamma / m)
    print("bn    C=%4d M=%6d dx rel=%.2e  dgamma rel=%.2e dbeta rel=%.2e" % (
        x.shape[1], m, rel(res[0], dx)[0], rel(res[1], dgamma)[0], rel(res[2], dbeta)[0]))
    return res


ops.BatchNormFn.backward = staticmethod(bn_bwd)

coords, feats, labels = synth.room_batch(777, 2, voxels)
model = models.Res16UNet34C(27, 20).to(dev).train()
field = ME.TensorField(coordinates=torch.from_numpy(coords).to(dev), features=torch.from_numpy(feats).to(dev))
out = model(field)
torch.nn.functional.cross_entropy(out, torch.from_numpy(labels).to(dev), ignore_index=255).backward()
