"""The other in-tree backbones on the GPU: MinkowskiPointNet against a torch fp64 restatement,
MinkowskiFCNN / MinkowskiSplatFCNN piecewise (strided slice, global pooling of a field) and end to end (shapes, finite
values, gradients on every parameter).  State-dict parity of these definitions with the reference's files:
tests/test_dropin.py."""
import numpy as np
import pytest
import torch

from nerf_downstream_b200 import me as ME
from nerf_downstream_b200 import models, ops, synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _field(coords, feats, dev):
    return ME.TensorField(coordinates=torch.from_numpy(coords).to(dev), features=torch.from_numpy(feats).to(dev))


def test_slice_from_strided_levels(cuda_device):
    """`y2.slice(x)` with y2 two stride-2 levels below the field (fcnn.py:162-165): every point reads the voxel
    floor(c / ts) * ts of that level."""
    coords, feats = synth.random_cloud(1, 3000, extent=9, n_batch=2, channels=4)
    x = _field(coords, feats, cuda_device)
    y = x.sparse()
    pool = ME.MinkowskiMaxPooling(kernel_size=3, stride=2, dimension=3)
    y2 = pool(y)
    y4 = pool(y2)
    for lvl, ts in ((y2, 2), (y4, 4)):
        got = lvl.slice(x).F.cpu().numpy()
        lut = {tuple(c): i for i, c in enumerate(lvl.C.cpu().numpy().tolist())}
        q = R.quantize_np(coords, (ts, ts, ts))
        rows = np.array([lut[tuple(c)] for c in q.tolist()])
        assert (got == lvl.F.detach().cpu().numpy()[rows]).all()
    # gradient of the strided slice lands on the voxel rows
    f = y2.F.detach().clone().requires_grad_(True)
    t = ME.SparseTensor(f, coordinate_map_key=y2.coordinate_map_key, coordinate_manager=y2.coordinate_manager)
    t.slice(x).F.sum().backward()
    counts = np.bincount(x.inverse_mapping(y2.coordinate_map_key).cpu().numpy(), minlength=f.shape[0])
    assert (f.grad[:, 0].cpu().numpy() == counts).all()


def test_global_pooling_of_a_field(cuda_device):
    coords, feats = synth.random_cloud(2, 2000, extent=5, n_batch=3, channels=6)
    x = _field(coords, feats, cuda_device)
    mx = ME.MinkowskiGlobalMaxPooling()(x)
    av = ME.MinkowskiGlobalAvgPooling()(x)
    assert mx.F.shape == (3, 6) and mx.C.cpu().tolist() == [[0, 0, 0, 0], [1, 0, 0, 0], [2, 0, 0, 0]]
    b = np.floor(coords[:, 0]).astype(int)
    for i in range(3):
        assert np.allclose(mx.F[i].cpu().numpy(), feats[b == i].max(0))
        assert np.allclose(av.F[i].cpu().numpy(), feats[b == i].mean(0), atol=1e-5)


def test_pointnet_matches_torch_restatement(cuda_device):
    torch.manual_seed(0)
    coords, feats = synth.random_cloud(3, 1500, extent=6, n_batch=3, channels=5)
    net = models.MinkowskiPointNet(5, 7, embedding_channel=64).to(cuda_device).train()
    net.dp1.module.p = 0.0                                         # dropout off: compare values
    sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
    logits = net(_field(coords, feats, cuda_device))
    logits.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())

    def bn(h, name):                                               # training-mode batch statistics, biased variance
        mu, var = h.mean(0), h.var(0, unbiased=False)
        return (h - mu) / torch.sqrt(var + 1e-5) * sd[f"{name}.1.bn.weight"] + sd[f"{name}.1.bn.bias"]
    h = torch.from_numpy(feats).double()
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        h = torch.relu(bn(h @ sd[f"{name}.0.linear.weight"].T, name))
    b = torch.from_numpy(np.floor(coords[:, 0]).astype(np.int64))
    g = torch.stack([h[b == i].max(0).values for i in range(3)])
    g = torch.relu(bn(g @ sd["linear1.0.linear.weight"].T, "linear1"))
    ref = g @ sd["linear2.linear.weight"].T + sd["linear2.linear.bias"]
    got = logits.detach().double().cpu()
    assert got.shape == (3, 7)
    assert float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)) >= 0.9999
    assert (got - ref).abs().max().item() <= 2e-3 * (1 + ref.abs().max().item())


@pytest.mark.parametrize("cls", [models.MinkowskiFCNN, models.MinkowskiSplatFCNN])
def test_fcnn_runs_end_to_end(cuda_device, cls):
    torch.manual_seed(1)
    coords, feats, labels = synth.co3d_batch(11, 3, channels=3, num_classes=10, lattice=40)
    net = cls(3, 10, embedding_channel=64, channels=(8, 16, 16, 32, 32)).to(cuda_device).train()
    logits = net(_field(coords, feats, cuda_device))
    assert logits.shape == (3, 10) and torch.isfinite(logits).all()
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels).to(cuda_device))
    loss.backward()
    missing = [n for n, p in net.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, missing
    assert sum(float(p.grad.abs().sum()) for p in net.parameters()) > 0
    # eval mode: same input -> same logits (up to the summation order of the split-offset convolutions)
    net.eval()
    with torch.no_grad():
        a = net(_field(coords, feats, cuda_device))
        b = net(_field(coords, feats, cuda_device))
    assert (a - b).abs().max().item() <= 1e-3 * (1 + a.abs().max().item())
