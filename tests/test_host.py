"""Host-side logic that runs without a GPU: C-ABI export surface, ME-compatible module
construction / state-dict keys, key & kernel-generator semantics, offsets, collate."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from nerf_downstream_b200 import build, lib
    path = build.build()
    header = (ROOT / "include" / "sparseconv_b200.h").read_text()
    declared = set(re.findall(r"\b(spc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    so = ctypes.CDLL(str(path))
    for name in sorted(declared):
        assert hasattr(so, name), f"{name} declared in the header but not exported"
    # the ctypes table mirrors the header one to one
    assert set(lib.SIGNATURES) == declared
    loaded = lib.load()
    assert loaded.spc_abi_version() == 1
    assert loaded.spc_table_slots(1000) == 2048 and loaded.spc_table_slots(1_000_000) == 2 ** 21


def test_no_cpu_fallback():
    from nerf_downstream_b200 import lib
    with pytest.raises(RuntimeError, match="CUDA"):
        lib.ptr(torch.zeros(4))
    import MinkowskiEngine as ME
    with pytest.raises(RuntimeError, match="no CPU backend"):
        ME.TensorField(coordinates=torch.zeros(3, 4), features=torch.zeros(3, 2))
    with pytest.raises(RuntimeError, match="no CPU backend"):
        ME.SparseTensor(torch.zeros(3, 2), coordinates=torch.zeros(3, 4, dtype=torch.int32))


def test_product_does_not_import_oracle():
    pkg = ROOT / "nerf_downstream_b200"
    for f in list(pkg.rglob("*.py")) + list((ROOT / "MinkowskiEngine").rglob("*.py")):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_conv_parameter_shapes_and_init():
    import MinkowskiEngine as ME
    torch.manual_seed(0)
    c = ME.MinkowskiConvolution(in_channels=27, out_channels=64, kernel_size=3, stride=1, dilation=1, bias=False,
                                dimension=3)
    assert tuple(c.kernel.shape) == (27, 27, 64) and c.bias is None and not c.use_mm
    assert c.kernel.abs().max() <= 1 / (27 * 27) ** 0.5 + 1e-7       # U(+-1/sqrt(Cin*K)), sparse_conv.py:427-435
    c = ME.MinkowskiConvolution(64, 64, kernel_size=1, stride=2, dimension=3)
    assert tuple(c.kernel.shape) == (1, 64, 64) and not c.use_mm        # k=1 s=2 is NOT the mm shortcut
    c = ME.MinkowskiConvolution(512, 51, kernel_size=1, bias=True, dimension=3)
    assert tuple(c.kernel.shape) == (512, 51) and c.use_mm and tuple(c.bias.shape) == (1, 51)
    t = ME.MinkowskiConvolutionTranspose(256, 128, kernel_size=2, stride=2, dimension=3)
    assert tuple(t.kernel.shape) == (8, 256, 128) and t.is_transpose
    assert t.kernel.abs().max() <= 1 / (128 * 8) ** 0.5 + 1e-7
    kg = c.kernel_generator
    assert kg.kernel_volume == 1 and kg.requires_strided_coordinates and kg.kernel_stride == [1, 1, 1]
    assert kg.region_type == ME.RegionType.HYPER_CUBE and kg.expand_coordinates is False


def test_state_dict_keys_match_reference_layout():
    import MinkowskiEngine as ME
    bn = ME.MinkowskiBatchNorm(8, momentum=0.05)
    assert list(bn.state_dict()) == ["bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
                                     "bn.num_batches_tracked"]
    assert bn.bn.momentum == 0.05 and isinstance(bn.bn, torch.nn.BatchNorm1d)
    lin = ME.MinkowskiLinear(4, 3)
    assert list(lin.state_dict()) == ["linear.weight", "linear.bias"]
    sync = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(bn))
    assert isinstance(sync[0], ME.MinkowskiSyncBatchNorm) and isinstance(sync[0].bn, torch.nn.SyncBatchNorm)


def test_model_parameter_counts():
    from nerf_downstream_b200 import models
    r = models.ResNet14(27, 51)
    n_conv = sum(p.numel() for n, p in r.named_parameters() if n.endswith("kernel") or n.endswith(".bias") and "bn" not in n)
    assert n_conv == 14_404_723
    u = models.Res16UNet34C(27, 20)
    n_conv = sum(p.numel() for n, p in u.named_parameters() if "bn" not in n)
    assert n_conv == 37_877_428
    keys = set(u.state_dict())
    for k in ["conv0p1s1.0.kernel", "conv0p1s1.4.bn.running_var", "block4.5.conv2.kernel",
              "block5.0.downsample.0.kernel", "convtr7p2s2.0.kernel", "final.bias", "block8.1.norm2.bn.weight"]:
        assert k in keys, k
    assert tuple(u.state_dict()["block5.0.downsample.0.kernel"].shape) == (384, 256)
    assert tuple(u.state_dict()["block8.0.conv1.kernel"].shape) == (27, 128, 96)


def test_keys_and_offsets():
    import MinkowskiEngine as ME
    from nerf_downstream_b200 import ops
    k = ME.CoordinateMapKey(4)
    assert k.get_coordinate_size() == 4 and not k.is_key_set()
    k.set_key([2, 2, 2], "")
    assert k == ME.CoordinateMapKey([2, 2, 2], "") and k != ME.CoordinateMapKey([2, 2, 2], "x")
    assert hash(k) == hash(ME.CoordinateMapKey([2, 2, 2]))
    offs = ops.kernel_offsets([3, 3, 3], [1, 1, 1], [1, 1, 1])
    assert offs[4] == (0, 0, -1) and offs[13] == (0, 0, 0) and offs[22] == (0, 0, 1)   # sparse_conv.py:375-379
    from oracle import ref_ops as R
    for ks, ts in [((3, 3, 3), (1, 1, 1)), ((3, 3, 3), (4, 4, 4)), ((2, 2, 2), (2, 2, 2)), ((1, 1, 1), (8, 8, 8))]:
        assert ops.kernel_offsets(ks, ts, (1, 1, 1)) == R.kernel_offsets(ks, ts)


def test_sparse_collate_float():
    import MinkowskiEngine as ME
    coords = [torch.tensor([[0.5, -1.5, 2.0]]), torch.tensor([[3.0, 4.0, 5.0], [6.0, 7.0, 8.5]])]
    feats = [torch.ones(1, 2), torch.zeros(2, 2)]
    c, f = ME.utils.sparse_collate(coords, feats, dtype=torch.float32)
    assert c.dtype == torch.float32 and c.tolist() == [[0, 0.5, -1.5, 2.0], [1, 3, 4, 5], [1, 6, 7, 8.5]]
    c, f, l = ME.utils.sparse_collate(coords, feats, [torch.tensor([1]), torch.tensor([2, 3])])
    assert c.dtype == torch.int32 and c.tolist() == [[0, 0, -2, 2], [1, 3, 4, 5], [1, 6, 7, 8]]
    assert l.tolist() == [1, 2, 3] and f.shape == (3, 2)


def test_faithful_geometry_generator_statistics():
    """the reference's real ScanNet-plenoxel transform puts samples ~2.3 voxels apart at the
    median scene_scale, so stride 1 has no collisions and (almost) only the centre tap, stride 2 removes nothing, and
    real neighbourhoods start at stride 2; a larger scene_scale packs them closer."""
    import numpy as np

    from nerf_downstream_b200 import synth
    from oracle import ref_ops as R
    stats = {}
    for scale in (0.34, 0.56):
        c, f, l = synth.faithful_room_batch(3, 2, 30_000, scene_scale=scale)
        assert c.shape == (60_000, 4) and f.shape == (60_000, 27) and l.shape == (60_000,) and c.dtype == np.float32
        assert set(np.unique(c[:, 0])) == {0.0, 1.0}
        uc, _, _ = R.unique_first_np(R.quantize_np(c))
        nbr = R.kernel_map_np(uc, uc, R.kernel_offsets((3, 3, 3), (1, 1, 1)))
        uc2 = R.stride_coords_np(uc, (2, 2, 2))
        nbr2 = R.kernel_map_np(uc2, uc2, R.kernel_offsets((3, 3, 3), (2, 2, 2)))
        stats[scale] = (uc.shape[0] / c.shape[0], (nbr >= 0).sum() / uc.shape[0], uc2.shape[0] / uc.shape[0],
                        (nbr2 >= 0).sum() / uc2.shape[0])
    m1, p1, m2, p2 = stats[0.34]
    assert m1 == 1.0 and p1 < 1.05 and m2 == 1.0 and p2 > 3.0
    assert stats[0.56][1] > 2.0 and stats[0.56][3] > p2
    pitch = 2 * 2 / (256 * 0.34 * 0.02)
    assert abs(pitch - 2.2978) < 1e-3
