"""Host logic of the ME-compatible layer on CPU, through tests/host_harness.py (READ ITS HEADER: the harness answers
the C-ABI calls with the oracle's numpy restatements so that the Python layer above the ABI — manager, key rules,
kernel-map caching, lazy BatchNorm fusion, channel padding, bf16 side copies, autograd wiring, training loop — can be
exercised without a GPU.  It says nothing about the CUDA kernels; tests/test_gpu_*.py cover those through the real
library).  The product itself still refuses CPU tensors (tests/test_host.py::test_no_cpu_fallback)."""
import numpy as np
import pytest
import torch

from nerf_downstream_b200 import ginlite, models, ops, pipeline, synth, training
from nerf_downstream_b200 import me as ME
from oracle import nets
from oracle import ref_ops as R
from tests import host_harness


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _field(coords, feats):
    return ME.TensorField(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats))


def _oracle_params(model):
    return {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "running" not in k)
            for k, v in model.state_dict().items()}


def test_resnet14_step_matches_oracle_network(monkeypatch):
    fake = host_harness.install(monkeypatch, "fp32")
    torch.manual_seed(0)
    coords, feats, labels = synth.co3d_batch(5, 2, lattice=24)
    model = models.ResNet14(27, 51).train()
    params = _oracle_params(model)
    y = torch.from_numpy(labels)
    ref = nets.resnet_forward(params, coords, torch.from_numpy(feats).double())
    torch.nn.functional.cross_entropy(ref, y).backward()
    out = model(_field(coords, feats))
    ops.cross_entropy(out, y).backward()
    assert _cos(out.detach(), ref.detach()) >= 0.999999
    for name, p in model.named_parameters():
        assert _cos(p.grad, params[name].grad) >= 0.9999, name
    # ResNet14 has one block per stage: every convolution has its own (map pair, kernel) -> one map each + the pool's
    n_maps = fake.calls.count("spc_kernel_map")
    n_convs = fake.calls.count("spc_conv_fwd")
    assert n_convs == 14 and n_maps <= n_convs + 1
    # BN + (residual) + ReLU run as ONE apply call per BatchNorm: no separate relu / add launches in the blocks
    assert fake.calls.count("spc_bn_apply") == 13 and fake.calls.count("spc_add") == 0
    assert fake.calls.count("spc_relu_fwd") == 0


def test_unet_fused_head_and_three_pass_agree_and_match_oracle(monkeypatch):
    host_harness.install(monkeypatch, "fp32")
    torch.manual_seed(1)
    coords, feats, labels = synth.room_batch(9, 2, 700, ignore_label=-255)
    model = models.Res16UNet14A(27, 20).train()
    params = _oracle_params(model)
    y = torch.from_numpy(labels)
    w = torch.ones(20)
    w[-1] = 0.3
    ref = nets.resunet_forward(params, coords, torch.from_numpy(feats).double())
    torch.nn.functional.cross_entropy(ref, y, weight=w.double(), ignore_index=-255).backward()

    logits = model(_field(coords, feats))
    loss_a = ops.cross_entropy(logits, y, -255, w)
    loss_a.backward()
    assert _cos(logits.detach(), ref.detach()) >= 0.99999
    grads_a = {n: p.grad.clone() for n, p in model.named_parameters()}
    for name, g in grads_a.items():
        assert _cos(g, params[name].grad) >= 0.999, name

    model.zero_grad()
    field = _field(coords, feats)
    counts = torch.zeros((3, 20), dtype=torch.int64)
    loss_b = pipeline.seg_head_loss(model.forward_sparse(field), field, y, -255, w, counts)
    loss_b.backward()
    assert abs(loss_a.item() - loss_b.item()) <= 1e-6 * abs(loss_a.item())
    for name, p in model.named_parameters():
        assert torch.allclose(p.grad, grads_a[name], rtol=1e-4, atol=1e-7), name
    assert (counts.numpy() == R.iou_counts_np(logits.detach().numpy(), labels, 20, -255)).all()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_tensor_core_modes_pad_channels_and_reuse_bf16_side_copies(monkeypatch, precision):
    """Large maps (>= 4096 rows) pad 27 -> 32 input and 20 -> 32 output channels for the tensor-core kernels and slice
    the padding off again; in bf16 mode BatchNorm hands its bf16 side copy to the next convolution (no conversion pass
    for those rows).  The harness computes in fp32 / bf16-rounded operands, so results must still match the oracle."""
    fake = host_harness.install(monkeypatch, precision)
    torch.manual_seed(2)
    coords, feats = synth.random_cloud(4, 9000, extent=14, n_batch=2, channels=27)
    x = _field(coords, feats).sparse()
    assert x.F.shape[0] >= 4096
    conv1 = ME.MinkowskiConvolution(27, 32, kernel_size=3, dimension=3)
    bn = ME.MinkowskiBatchNorm(32)
    relu = ME.MinkowskiReLU()
    conv2 = ME.MinkowskiConvolution(32, 20, kernel_size=3, bias=True, dimension=3)
    out = conv2(relu(bn(conv1(x))))
    out.F.sum().backward()
    assert out.F.shape == (x.F.shape[0], 20) and conv1.kernel.grad.shape == (27, 27, 32)
    assert conv2.kernel.grad.shape == (27, 32, 20) and conv2.bias.grad.shape == (1, 20)
    # oracle of the same two layers
    uc = x.C.numpy()
    nbr = R.kernel_map_np(uc, uc, R.kernel_offsets((3, 3, 3), (1, 1, 1)))
    h = R.conv_forward(x.F.detach().double(), conv1.kernel.detach().double(), nbr)
    h = torch.relu(R.batch_norm(h, bn.bn.weight.detach().double(), bn.bn.bias.detach().double()))
    ref = R.conv_forward(h, conv2.kernel.detach().double(), nbr, conv2.bias.detach().double())
    tol = 1e-5 if precision == "tf32" else 3e-2                  # the harness rounds bf16 operands like the kernels
    assert (out.F.detach().double() - ref).abs().max().item() <= tol * ref.abs().max().item()
    conversions = fake.calls.count("spc_to_bf16")
    if precision == "bf16":
        # forward: input rows of conv1 only (conv2 takes BatchNorm's side copy); backward: dout of conv2, and the
        # gradient that reaches conv1 comes out of BatchNorm backward with its bf16 copy attached
        assert conversions == 2, fake.calls
    else:
        assert conversions == 0
    assert fake.calls.count("spc_tile_mask") >= 1                # tensor-core paths ask for the per-tile offset masks


def test_manager_key_rules_and_caching(monkeypatch):
    fake = host_harness.install(monkeypatch, "fp32")
    coords, feats = synth.random_cloud(3, 1500, extent=8, n_batch=2, channels=8)
    x = _field(coords, feats).sparse()
    mgr = x.coordinate_manager
    a = ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=2, dimension=3)(x)
    b = ME.MinkowskiConvolution(8, 8, kernel_size=1, stride=2, dimension=3)(x)
    assert a.coordinate_map_key == b.coordinate_map_key and a.tensor_stride == [2, 2, 2]
    assert fake.calls.count("spc_coords_insert") == 2            # field -> voxels, stride-2 map: built once
    up = ME.MinkowskiConvolutionTranspose(8, 4, kernel_size=2, stride=2, dimension=3)(a)
    assert up.coordinate_map_key == x.coordinate_map_key
    n_maps = fake.calls.count("spc_kernel_map")
    ME.MinkowskiConvolution(8, 8, kernel_size=3, stride=2, dimension=3)(x)       # same (keys, kernel): cached
    assert fake.calls.count("spc_kernel_map") == n_maps
    with pytest.raises(AssertionError):
        ME.cat(a, x)
    pairs = mgr.kernel_map(x.coordinate_map_key, x.coordinate_map_key, 1, 3, 1)
    assert bool((pairs[13][0] == pairs[13][1]).all()) and pairs[13].shape[1] == x.F.shape[0]
    g = ME.MinkowskiGlobalAvgPooling()(a)
    assert g.F.shape == (2, 8) and g.C.tolist() == [[0, 0, 0, 0], [1, 0, 0, 0]]
    # out-of-range coordinates surface as a RuntimeError, as from the real library
    far = coords.copy()
    far[0, 1] = 3.0e5
    with pytest.raises(RuntimeError, match="out of the supported range"):
        _field(far, feats).sparse()


@pytest.mark.parametrize("cls", [models.MinkowskiFCNN, models.MinkowskiSplatFCNN])
def test_fcnn_backbones_run_on_the_surface(monkeypatch, cls):
    host_harness.install(monkeypatch, "fp32")
    torch.manual_seed(3)
    coords, feats, labels = synth.co3d_batch(11, 2, channels=3, num_classes=10, lattice=20)
    net = cls(3, 10, embedding_channel=32, channels=(4, 8, 8, 8, 16)).train()
    logits = net(_field(coords, feats))
    assert logits.shape == (2, 10) and torch.isfinite(logits).all()
    torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_pointnet_matches_torch_restatement(monkeypatch):
    host_harness.install(monkeypatch, "fp32")
    torch.manual_seed(4)
    coords, feats = synth.random_cloud(3, 600, extent=6, n_batch=3, channels=5)
    net = models.MinkowskiPointNet(5, 7, embedding_channel=32).train()
    net.dp1.module.p = 0.0
    sd = {k: v.detach().double() for k, v in net.state_dict().items()}
    logits = net(_field(coords, feats))

    def bn(h, name):
        return (h - h.mean(0)) / torch.sqrt(h.var(0, unbiased=False) + 1e-5) * sd[f"{name}.1.bn.weight"] + sd[f"{name}.1.bn.bias"]
    h = torch.from_numpy(feats).double()
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        h = torch.relu(bn(h @ sd[f"{name}.0.linear.weight"].T, name))
    b = torch.from_numpy(np.floor(coords[:, 0]).astype(np.int64))
    g = torch.stack([h[b == i].max(0).values for i in range(3)])
    g = torch.relu(bn(g @ sd["linear1.0.linear.weight"].T, "linear1"))
    ref = g @ sd["linear2.linear.weight"].T + sd["linear2.linear.bias"]
    assert _cos(logits.detach(), ref) >= 0.99999


def test_gin_driven_run_on_the_surface(monkeypatch, tmp_path):
    """`training.train` end to end (gin file -> get_model -> fit with the fused head -> Lightning checkpoint ->
    evaluate) with the real `models.Res16UNet14A` on the ME surface."""
    host_harness.install(monkeypatch, "fp32")
    ginlite.clear_config()
    try:
        cfg = tmp_path / "tiny.gin"
        cfg.write_text('get_model.name = "Res16UNet14A"\nget_model.in_channel = 27\nget_model.out_channel = 20\n'
                       'train.max_steps = 3\ntrain.scheduler_name = "PolyLR"\nPolyLR.poly_exp = 0.9\ntrain.lr = 0.05\n'
                       'train.ignore_label = -255\ntrain.val_every_n_steps = 3\ntrain.log_every_n_steps = 1\n'
                       'SGD.momentum = 0.9\n')
        torch.manual_seed(5)

        def batches(seed, k):
            out = []
            for i in range(k):
                c, f, y = synth.room_batch(seed + i, 1, 400, ignore_label=-255)
                out.append({"coordinates": torch.from_numpy(c), "features": torch.from_numpy(f), "labels": torch.from_numpy(y)})
            return out
        tb, vb = batches(10, 2), batches(20, 1)
        logs = []
        run = training.train([str(cfg)], [], lambda: tb, lambda: vb, save_path=str(tmp_path / "run"), device="cpu",
                             log=logs.append, fused_head=True)
        assert run.global_step == 3 and [d["global_step"] for d in logs if "train/loss" in d] == [1, 2]
        last = [d for d in logs if "val/mIoU" in d][-1]
        res = training.evaluate(str(tmp_path / "run" / "last.ckpt"), vb, model=models.Res16UNet14A(27, 20), tag="e",
                                fused_head=True)
        assert abs(res["val/mIoU"] - last["val/mIoU"]) < 1e-3 and abs(res["val/loss"] - last["val/loss"]) < 1e-4
    finally:
        ginlite.clear_config()


def test_stride_pyramid_equals_level_by_level(monkeypatch):
    """The first stride-2 request builds the 2-4-8-16 maps back to back, each level reading its parent's row count
    from the parent's status word (one host synchronisation for all levels): maps, parent rows and the network
    output must equal the level-by-level construction."""
    outs, sizes, calls = [], [], []
    for depth in (1, 4):
        fake = host_harness.install(monkeypatch, "fp32", prefetch_depth=depth)
        torch.manual_seed(7)
        coords, feats, _ = synth.room_batch(13, 2, 600)
        model = models.Res16UNet14A(27, 20).train()
        field = _field(coords, feats)
        outs.append(model(field).detach())
        mgr = field.coordinate_manager
        sizes.append({tuple(k.get_tensor_stride()): (mgr.size(k), mgr.get_coordinates(k).clone())
                      for k in mgr._maps if k.get_tensor_stride()[0] in (2, 4, 8, 16)})
        calls.append((fake.calls.count("spc_coords_insert"), fake.calls.count("spc_coords_insert_dev")))
    assert calls[0][1] == 0 and calls[1][1] == 4                        # four strided levels in one go
    assert calls[1][0] == calls[0][0] - 4
    assert sizes[0].keys() == sizes[1].keys() and len(sizes[0]) == 4
    for k in sizes[0]:
        assert sizes[0][k][0] == sizes[1][k][0] and torch.equal(sizes[0][k][1], sizes[1][k][1]), k
    assert torch.equal(outs[0], outs[1])


def test_record_to_sample_matches_reference_getitem(monkeypatch):
    """`pipeline.co3d_sample` / `scannet_plenoxel_sample` against the reference's own `Dataset.__getitem__` code run on
    the same in-memory plenoxel record (tests/golden/make_samples.py -> samples_ref.npz): decode, void labelling,
    lattice thinning, normalisation, the train transformations under the same seeds (parameters from gin bindings),
    feature selection, NYU40 -> 20-class label map.  (Decode kernel emulated by the harness; its GPU parity is
    tests/test_gpu_parity.py::test_plenoxel_decode_exact.)"""
    import random
    from pathlib import Path
    host_harness.install(monkeypatch, "fp32")
    gold = Path(__file__).resolve().parent / "golden"
    ref = np.load(gold / "samples_ref.npz")
    rec = {k.split("/", 1)[1]: ref[k] for k in ref.files if k.startswith("record/")}
    links, density = torch.from_numpy(rec["links"]), torch.from_numpy(rec["density"])
    sh = torch.from_numpy(rec["sh"])
    scale, mn, reso = float(rec["sh_scale"]), float(rec["sh_min"]), [int(v) for v in rec["reso"]]
    for case, feats in (("co3d_plain", ["sh"]), ("co3d_xyz_density", ["xyzs", "density", "ones"])):
        s = pipeline.co3d_sample(links, density, sh, scale, mn, reso, features=feats)
        assert (s["coordinates"].numpy() == ref[f"{case}/coordinates"]).all()
        assert np.abs(s["features"].numpy() - ref[f"{case}/features"]).max() <= 1e-6, case
        assert np.abs(s["xyzs"].numpy() - ref[f"{case}/xyzs"]).max() <= 1e-6
    ginlite.clear_config()
    try:
        ginlite.parse_config((gold / "samples_ref.gin").read_text())
        labels, dists = torch.from_numpy(rec["labels"]), torch.from_numpy(rec["dists"])
        names = ["RandomRotation", "RandomCrop", "RandomAffine", "CoordinateDropout", "RandomHorizontalFlip",
                 "RandomTranslation", "ElasticDistortion"]
        for case, feats, tf, seed, void in (("scannet_plain", ["sh"], [], 0, None),
                                            ("scannet_void", ["density", "sh"], [], 0, 40),
                                            ("scannet_aug", ["sh"], names, 2, None)):
            random.seed(seed)
            np.random.seed(seed)
            s = pipeline.scannet_plenoxel_sample(links, density, sh, scale, mn, reso, labels, dists, scene_scale=0.34,
                                                 ignore_label=-255, void_label=void, features=feats, transformations=tf)
            want = ref[f"{case}/coordinates"]
            assert s["coordinates"].shape == want.shape, (case, s["coordinates"].shape, want.shape)
            assert (s["labels"].numpy() == ref[f"{case}/labels"]).all(), case
            tol = 1e-4 if tf else 2e-5                           # float32 chain vs the reference's float64 chain
            assert np.abs(s["coordinates"].numpy() - want).max() <= tol * max(1.0, np.abs(want).max()), case
            assert np.abs(s["features"].numpy() - ref[f"{case}/features"]).max() <= 1e-6
            assert np.abs(s["dists"].numpy() - ref[f"{case}/dists"]).max() <= 1e-7
        assert (ref["scannet_void/labels"] == 20).any() and (ref["scannet_plain/labels"] == -255).any()
    finally:
        ginlite.clear_config()
    with pytest.raises(KeyError, match="unknown feature"):
        pipeline.co3d_sample(links, density, sh, scale, mn, reso, features=["rgb"])


def test_surface_inventory_of_survey_8b(monkeypatch):
    """Every MinkowskiEngine symbol / attribute / signature the reference touches (SURVEY.md §8b "Symbols + signatures
    actually used") exists and behaves: tensors, layers, functional forms, manager, keys, kernel generator, utils."""
    import MinkowskiEngine as MEpkg
    import MinkowskiEngine.MinkowskiFunctional as MEF
    from MinkowskiEngine.MinkowskiCoordinateManager import CoordinateManager
    from MinkowskiEngine.MinkowskiKernelGenerator import KernelGenerator
    from MinkowskiEngine.MinkowskiSparseTensor import CoordinateMapKey, SparseTensor
    from MinkowskiEngineBackend._C import ConvolutionMode, RegionType  # noqa: F401
    host_harness.install(monkeypatch, "fp32")
    coords, feats = synth.random_cloud(5, 900, extent=7, n_batch=2, channels=6)
    f = MEpkg.TensorField(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats))
    assert f.F.shape == (900, 6) and f.device == torch.device("cpu") and f.dtype == torch.float32
    assert f.quantization_mode == MEpkg.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE
    f2 = MEpkg.TensorField(f.F * 2, coordinate_field_map_key=f.coordinate_field_map_key,
                           coordinate_manager=f.coordinate_manager, quantization_mode=f.quantization_mode)   # layernorm.py:19-24
    x = f.sparse()
    x2 = f2.sparse(quantization_mode=MEpkg.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE)
    assert isinstance(x, SparseTensor) and x.D == 3 and x.shape == x.F.shape and x.tensor_stride == [1, 1, 1]
    assert x._manager is x.coordinate_manager and isinstance(x.coordinate_manager, CoordinateManager)
    assert torch.allclose(x2.F, 2 * x.F)
    y = MEpkg.SparseTensor(x.F + 1, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)
    y += x                                                                            # resnet_block.py:66
    assert torch.allclose(y.F, 2 * x.F + 1) and y.slice(f).F.shape == (900, 6)
    s = MEpkg.SparseTensor(features=torch.from_numpy(feats), coordinates=MEpkg.utils.batched_coordinates(
        [torch.from_numpy(coords[:, 1:])], dtype=torch.float32), device="cpu")       # co3d.py:119, transforms.py:508-512
    assert s.F.shape[1] == 6 and s.C.dtype == torch.int32
    # layers + functional forms
    conv = MEpkg.MinkowskiConvolution(in_channels=6, out_channels=8, kernel_size=3, stride=1, dilation=1, bias=True, dimension=3)
    assert conv.in_channels == 6 and conv.out_channels == 8 and tuple(conv.kernel.shape) == (27, 6, 8) and conv.bias.shape == (1, 8)
    kg = conv.kernel_generator
    assert (kg.kernel_size, kg.kernel_stride, kg.kernel_dilation, kg.kernel_volume) == ([3, 3, 3], [1, 1, 1], [1, 1, 1], 27)
    assert kg.region_type == RegionType.HYPER_CUBE and kg.expand_coordinates is False
    assert isinstance(kg.requires_strided_coordinates, bool)
    h = conv(x)
    for name in ("MinkowskiReLU", "MinkowskiPReLU", "MinkowskiLeakyReLU", "MinkowskiELU", "MinkowskiCELU", "MinkowskiSELU",
                 "MinkowskiGELU"):
        assert getattr(MEpkg, name)()(h).F.shape == h.F.shape, name
    assert MEpkg.MinkowskiReLU(inplace=True)(h).F.min() >= 0
    for name in ("relu", "leaky_relu", "prelu", "celu", "selu", "gelu"):
        fn = getattr(MEF, name)
        out = fn(h, torch.tensor([0.25])) if name == "prelu" else fn(h)
        assert out.F.shape == h.F.shape and out.coordinate_map_key == h.coordinate_map_key, name
    assert MEpkg.MinkowskiSumPooling(kernel_size=2, stride=2, dimension=3)(h).tensor_stride == [2, 2, 2]
    avg = MEpkg.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(h)                     # co3d.py:107-111
    assert avg.C[:, 1:].float().shape[1] == 3 and avg.F.shape[1] == 8
    assert MEpkg.MinkowskiGlobalAvgPooling()(h).F.shape == (2, 8)
    lin = MEpkg.MinkowskiLinear(8, 4, bias=False)
    assert isinstance(lin.linear, torch.nn.Linear) and lin(h).F.shape[1] == 4
    bn = MEpkg.MinkowskiBatchNorm(8, momentum=0.05)
    assert isinstance(bn.bn, torch.nn.BatchNorm1d) and bn.bn.momentum == 0.05
    assert isinstance(MEpkg.MinkowskiNetwork(3), torch.nn.Module) and issubclass(MEpkg.MinkowskiConvolution, MEpkg.MinkowskiModuleBase)
    net = torch.nn.Sequential(conv, bn)
    synced = MEpkg.MinkowskiSyncBatchNorm.convert_sync_batchnorm(net)                              # train.py:106-107
    assert isinstance(synced[1], MEpkg.MinkowskiSyncBatchNorm) and synced[1].bn.weight is bn.bn.weight
    assert MEpkg.cat(h, h).F.shape[1] == 16
    # manager + keys (sparse_conv.py:80-96,397-405)
    cm = x.coordinate_manager
    key = CoordinateMapKey(x.coordinate_map_key.get_coordinate_size())
    assert not key.is_key_set()
    key.set_key([2, 2, 2], "")
    assert cm.stride(x.coordinate_map_key, [2, 2, 2]) == key and cm.size(key) > 0
    pairs = cm.kernel_map(x.coordinate_map_key, key, [2, 2, 2], [3, 3, 3], [1, 1, 1], is_transpose=False)
    assert all(p.shape[0] == 2 and p.dtype == torch.int32 for p in pairs.values())
    assert KernelGenerator(kernel_size=2, stride=2, dilation=1, expand_coordinates=False, dimension=3).kernel_volume == 8
    # utils
    cb, fb, lb = MEpkg.utils.sparse_collate([coords[:10, 1:], coords[10:30, 1:]], [feats[:10], feats[10:30]],
                                            [np.zeros(10), np.ones(20)], dtype=torch.float32)
    assert cb.shape == (30, 4) and cb.dtype == torch.float32 and fb.shape == (30, 6) and lb.shape[0] == 30
    w = torch.empty(27, 6, 8)
    MEpkg.utils.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
    assert abs(float(w.std()) - (2.0 / (27 * 8)) ** 0.5) < 0.02


def test_torch_prune_on_minkowski_convolution(monkeypatch):
    """The reference prunes `MinkowskiConvolution.kernel` with torch.nn.utils.prune (utils/prune.py:26-76): the
    re-parametrised module (`kernel_orig` + `kernel_mask`) still runs, equals the convolution with the masked kernel,
    trains only the kept weights, and `count_parameters` reports the removed ones."""
    import torch.nn.utils.prune as torch_prune
    host_harness.install(monkeypatch, "fp32")
    coords, feats = synth.random_cloud(8, 700, extent=6, n_batch=2, channels=8)
    x = _field(coords, feats).sparse()
    torch.manual_seed(0)
    net = torch.nn.Sequential(ME.MinkowskiConvolution(8, 16, kernel_size=3, dimension=3), ME.MinkowskiBatchNorm(16),
                              ME.MinkowskiReLU(), ME.MinkowskiConvolution(16, 4, kernel_size=1, bias=True, dimension=3))
    dense = net[0].kernel.detach().clone()
    torch_prune.l1_unstructured(net[0], "kernel", amount=0.5)
    assert {"kernel_orig"} <= {n for n, _ in net[0].named_parameters()} and "kernel_mask" in dict(net[0].named_buffers())
    out = net(x)
    ref_conv = ME.MinkowskiConvolution(8, 16, kernel_size=3, dimension=3)
    with torch.no_grad():
        ref_conv.kernel.copy_(dense * net[0].kernel_mask)
    assert torch.allclose(net[0](x).F, ref_conv(x).F, atol=1e-6)
    out.F.sum().backward()
    g = net[0].kernel_orig.grad
    assert g is not None and float(g[net[0].kernel_mask == 0].abs().max()) == 0.0 and float(g.abs().max()) > 0
    counts = training.count_parameters(net)
    assert counts["pruned"] == float((net[0].kernel_mask == 0).sum()) == 27 * 8 * 16 // 2
    assert counts["total"] == float(sum(p.numel() for p in net.parameters()))


def test_evaluate_pruned_checkpoint(monkeypatch, tmp_path):
    """eval.py:47-74: a Lightning checkpoint of a network pruned with torch.nn.utils.prune (`kernel_orig` /
    `kernel_mask` entries) evaluates in a fresh model to the pruned network's own numbers."""
    import torch.nn.utils.prune as torch_prune
    host_harness.install(monkeypatch, "fp32")
    ginlite.clear_config()
    try:
        ginlite.parse_config("get_model.name = 'Res16UNet14A'\nget_model.in_channel = 27\nget_model.out_channel = 20\n"
                             "train.ignore_label = -255")
        torch.manual_seed(3)
        net = models.Res16UNet14A(27, 20)
        to_prune = training.get_parameters_to_prune(net)
        assert len(to_prune) > 20 and all(name == "kernel" for _, name in to_prune)
        torch_prune.global_unstructured(to_prune, pruning_method=torch_prune.L1Unstructured, amount=0.4)
        sd = training.lightning_state_dict(net)
        assert any(k.endswith("kernel_mask") for k in sd) and any(k.endswith("kernel_orig") for k in sd)
        torch.save({"state_dict": sd, "global_step": 7}, tmp_path / "pruned.ckpt")
        c, f, y = synth.room_batch(4, 1, 400, ignore_label=-255)
        val = [{"coordinates": torch.from_numpy(c), "features": torch.from_numpy(f), "labels": torch.from_numpy(y)}]
        res = training.evaluate(str(tmp_path / "pruned.ckpt"), val, model=models.Res16UNet14A(27, 20), tag="p")
        want = training.Run(net, training.TrainConfig(max_steps=0, ignore_label=-255), evaluate_only=True).validate(val)
        assert abs(res["val/mIoU"] - want["val/mIoU"]) < 1e-3 and abs(res["val/loss"] - want["val/loss"]) < 1e-4
        total = sum(p.numel() for _, p in net.named_parameters())
        assert res["val/total_params"] == float(total) and abs(res["val/pruned_params"] / sum(m.kernel_mask.numel() for m, _ in to_prune) - 0.4) < 1e-3
    finally:
        ginlite.clear_config()


def _two_blocks(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(models.ResidualBlock(32, 32), models.ResidualBlock(32, 32))


def _run_blocks(monkeypatch, fused: bool):
    """stem conv (27 -> 32, padded: not fusable) + BN + ReLU, two residual blocks, 1x1 head; returns
    (logits, {parameter: gradient}, call list)."""
    from nerf_downstream_b200 import ops as O
    fake = host_harness.install(monkeypatch, "bf16")
    for knob in ("fuse_conv_bn", "hollow_rows", "recompute_relu_mask", "fuse_residual_grad", "fuse_bn_stats"):
        monkeypatch.setattr(O, knob, fused)
    torch.manual_seed(5)
    coords, feats = synth.random_cloud(6, 9000, extent=14, n_batch=2, channels=27)
    stem = torch.nn.Sequential(ME.MinkowskiConvolution(27, 32, kernel_size=3, dimension=3), ME.MinkowskiBatchNorm(32),
                               ME.MinkowskiReLU())
    blocks = _two_blocks(6)
    head = ME.MinkowskiConvolution(32, 20, kernel_size=1, bias=True, dimension=3)
    net = torch.nn.Sequential(stem, blocks, head).train()
    out = net(_field(coords, feats).sparse())
    logits = out.F
    (logits * torch.linspace(-1, 1, 20)).sum().backward()
    return logits.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}, list(fake.calls), net


def test_fused_conv_bn_node_hollow_rows_and_recomputed_relu_mask_agree_with_the_separate_nodes(monkeypatch):
    """bf16 mode: conv + BN (+ residual, ReLU) as ONE autograd node, BatchNorm outputs that only feed convolutions
    without fp32 rows, ReLU masks re-computed from x — same logits and parameter gradients as conv / BN as separate
    nodes with every tensor materialised (the harness rounds bf16 operands like the kernels do)."""
    from nerf_downstream_b200 import ops as O
    made0, acc0, add0 = O.hollow_stats["made"], O.residual_stats["accumulated"], O.residual_stats["added"]
    la, ga, calls_a, net_a = _run_blocks(monkeypatch, True)
    made = O.hollow_stats["made"] - made0
    # residual gradients: both blocks' conv1 dgrad reduce-add into the gradient BatchNorm backward wrote for the
    # residual branch (no separate sum of two activation gradients)
    assert O.residual_stats["accumulated"] - acc0 == 2 and O.residual_stats["added"] == add0
    lb, gb, calls_b, net_b = _run_blocks(monkeypatch, False)
    assert torch.allclose(la, lb, rtol=1e-5, atol=1e-6)
    for name, g in ga.items():
        assert torch.allclose(g, gb[name], rtol=2e-4, atol=1e-6), name
    # running statistics advanced once per BatchNorm in both
    for (na, ba), (nb, bb) in zip(net_a.named_buffers(), net_b.named_buffers()):
        assert torch.allclose(ba.float(), bb.float(), rtol=1e-5, atol=1e-7), na
    # the stem's BN output and bn1 of each block feed one convolution only: hollow; the stem output is then read as
    # block 1's residual and filled (one extra apply call); block outputs (residual folded in) are never hollow
    assert made == 3
    assert calls_a.count("spc_bn_apply") == calls_b.count("spc_bn_apply") + 1
    # one launch set per layer either way: the fused node adds no kernels
    for name in ("spc_conv_fwd_packed", "spc_conv_dgrad_packed", "spc_bn_bwd"):
        assert calls_a.count(name) == calls_b.count(name), name
    # the four fused nodes take their BatchNorm statistics from the convolution's epilogue (spc_bn_finalize); the
    # stem (padded channels: separate nodes) still runs the statistics pass
    assert calls_a.count("spc_bn_finalize") == 4 and calls_a.count("spc_bn_stats") == 1
    assert calls_b.count("spc_bn_finalize") == 0 and calls_b.count("spc_bn_stats") == 5
    # gradients of the conv outputs inside the fused nodes are bf16-only: fewer conversion passes, never more
    assert calls_a.count("spc_to_bf16") <= calls_b.count("spc_to_bf16")


def test_pending_rows_keep_the_grad_mode_and_training_flag_of_the_module_call(monkeypatch):
    host_harness.install(monkeypatch, "fp32")
    coords, feats = synth.random_cloud(7, 800, extent=6, n_batch=1, channels=8)
    x = _field(coords, feats).sparse()
    conv, bn = ME.MinkowskiConvolution(8, 8, kernel_size=3, dimension=3), ME.MinkowskiBatchNorm(8)
    with torch.no_grad():
        y = bn(conv(x))
    assert not y.F.requires_grad                    # produced outside the no_grad block, but called inside it
    bn.eval()
    z = bn(conv(x))
    bn.train()
    tracked = int(bn.bn.num_batches_tracked)
    z.F                                             # an eval-mode call: running statistics, no update
    assert int(bn.bn.num_batches_tracked) == tracked
    w = ME.MinkowskiReLU()(bn(conv(x)))
    with pytest.raises(RuntimeError, match="pre-activation"):
        _ = bn(conv(x)), None
        pre = bn(conv(x))
        ME.MinkowskiReLU()(pre)
        pre.F
    assert w.F.min().item() >= 0


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_dgrad_of_a_symmetric_self_map_reads_the_forward_map_with_reversed_offsets(monkeypatch, precision):
    """nbr_t[k] == nbr[K-1-k] on a centrally symmetric self map: dgrad with the reversed-offset weight image on the
    forward map equals dgrad on the transposed map, and the transposed map is never built."""
    from nerf_downstream_b200 import ops as O
    grads, calls = [], []
    for sym in (True, False):
        fake = host_harness.install(monkeypatch, precision)
        monkeypatch.setattr(O, "symmetric_dgrad", sym)
        monkeypatch.setattr(O, "_pack_cache", {})
        torch.manual_seed(8)
        coords, feats = synth.random_cloud(9, 9000, extent=14, n_batch=2, channels=32)
        x = _field(coords, feats).sparse()
        f = x.F.detach().requires_grad_(True)
        xs = ME.SparseTensor(f, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)
        conv = ME.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3)
        out = conv(xs).F
        (out * torch.linspace(-1, 1, 64)).sum().backward()
        grads.append(f.grad.clone())
        calls.append(list(fake.calls))
    assert torch.allclose(grads[0], grads[1], rtol=1e-5, atol=1e-6)
    assert calls[0].count("spc_kernel_map_transpose") == 0 and calls[1].count("spc_kernel_map_transpose") == 1


def test_unet_in_bf16_mode_with_lazy_cat_and_fused_nodes_matches_the_plain_graph(monkeypatch):
    """Res16UNet14A on the host harness in bf16 mode: ME.cat assembled from the bf16 operand copies (ops.CatFn, fp32
    concatenation hollow), conv + BN nodes fused, hollow BatchNorm rows, re-computed ReLU masks, symmetric dgrad — the
    logits and every parameter gradient equal those of the plain graph (torch.cat, separate nodes, all rows written)."""
    from nerf_downstream_b200 import ops as O
    res = []
    for on in (True, False):
        fake = host_harness.install(monkeypatch, "bf16")
        for knob in ("fuse_conv_bn", "hollow_rows", "recompute_relu_mask", "lazy_cat", "symmetric_dgrad",
                     "fuse_residual_grad", "fuse_bn_stats"):
            monkeypatch.setattr(O, knob, on)
        monkeypatch.setattr(O, "_pack_cache", {})
        torch.manual_seed(12)
        coords, feats = synth.random_cloud(13, 6000, extent=12, n_batch=2, channels=27)
        labels = torch.from_numpy(np.random.RandomState(0).randint(0, 20, size=coords.shape[0]))
        net = models.Res16UNet14A(27, 20).train()
        out = net(_field(coords, feats))
        loss = ops.cross_entropy(out, labels, -255)
        loss.backward()
        res.append((out.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}, list(fake.calls)))
    (la, ga, ca), (lb, gb, cb) = res
    assert torch.allclose(la, lb, rtol=1e-4, atol=1e-5)
    for n in ga:
        assert _cos(ga[n], gb[n]) >= 0.9999, n
    assert ca.count("spc_copy_rows") == 8 and cb.count("spc_copy_rows") == 0      # 4 skip connections x 2 parts
    assert ca.count("spc_kernel_map_transpose") < cb.count("spc_kernel_map_transpose")
    assert ca.count("spc_to_bf16") < cb.count("spc_to_bf16")


def test_hollow_registry_does_not_keep_tensors_alive(monkeypatch):
    """A hollow tensor's registry entry (and the rows its fill closure needs) must die with the tensor."""
    import gc
    from nerf_downstream_b200 import ops as O
    host_harness.install(monkeypatch, "bf16")
    x = torch.randn(64, 32)
    bn = torch.nn.BatchNorm1d(32)
    y = O.BatchNormFn.apply(x, bn.weight, bn.bias, None, None, True, 0.1, 1e-5, True, None, None, False)
    assert O.is_hollow(y)
    n = len(O._hollow)
    del y
    gc.collect()
    assert len(O._hollow) == n - 1
    cat = O.CatFn.apply(torch.randn(16, 32), torch.randn(16, 32))
    assert O.is_hollow(cat)
    n = len(O._hollow)
    del cat
    gc.collect()
    assert len(O._hollow) == n - 1


def test_row_selections_of_a_plenoxel_record_follow_the_reference_transforms(monkeypatch):
    """RandomCrop / CoordinateDropout as a row list + one decode of the kept records == decoding everything and pushing
    the tensors through the (reference-pinned) transforms of augment.py with the same RNG state."""
    import random as py_random

    from nerf_downstream_b200 import augment
    fake = host_harness.install(monkeypatch, "fp32")
    rng = np.random.RandomState(3)
    reso = (64, 48, 40)
    links = torch.from_numpy(np.sort(rng.choice(reso[0] * reso[1] * reso[2], 5000, replace=False)).astype(np.int32))
    sh = torch.from_numpy(rng.randint(0, 256, size=(5000, 27)).astype(np.uint8))
    aff = [1.0, 0.05, 0, -0.05, 1.0, 0, 0, 0, 1.0, 2.0, -1.0, 0.5]
    steps = [("RandomCrop", dict(x=30, y=25, z=100, application_ratio=1.0)),
             ("CoordinateDropout", dict(dropout_ratio=0.3, application_ratio=1.0))]
    py_random.seed(7)
    np.random.seed(7)
    rows = pipeline.plenoxel_select_rows(links, reso, steps, aff)
    assert rows is not None and rows.dtype == torch.int32
    c_a, f_a = pipeline.plenoxel_decode(links, sh, 2.0 / 255, -1.0, reso, affine=aff, rows=rows)
    # the same through the transforms on fully decoded tensors
    py_random.seed(7)
    np.random.seed(7)
    c_all, f_all = pipeline.plenoxel_decode(links, sh, 2.0 / 255, -1.0, reso, affine=aff)
    xyz, f_b, _ = augment.random_crop(c_all[:, 1:], f_all, None, 30, 25, 100, 1.0)
    xyz, f_b, _ = augment.coordinate_dropout(xyz, f_b, None, 0.3, 1.0)
    assert 0 < xyz.shape[0] < 5000
    assert torch.equal(c_a[:, 1:], xyz) and torch.equal(f_a, f_b)
    assert fake.calls.count("spc_plenoxel_decode_rows") == 1 and fake.calls.count("spc_plenoxel_crop_select") >= 1
    # a box larger than the extent: the reference returns its input and draws nothing
    state = np.random.get_state()[1].copy()
    assert pipeline.plenoxel_select_rows(links, reso, [("RandomCrop", dict(x=500, y=500, z=500))], aff) is None
    assert (np.random.get_state()[1] == state).all()


def test_engine_side_row_order_leaves_results_at_the_points_unchanged(monkeypatch):
    """ops.reorder_rows_by_mask on the host layer (harness): a sparse cloud (few neighbours per voxel, every raster tile
    touching most offsets) is re-ordered, strided levels follow their parents' new numbering, and the network's output
    at the points of the TensorField, the loss and the parameter gradients equal those of the first-occurrence order."""
    def run(sort):
        fake = host_harness.install(monkeypatch, "fp32", prefetch_depth=4)
        monkeypatch.setattr(ops, "sort_rows", sort)
        monkeypatch.setattr(ops, "sort_min_rows", 512)
        monkeypatch.setattr(ops, "sort_window", 1024)
        monkeypatch.setattr(ops, "sort_min_ratio", 1.2)
        monkeypatch.setattr(ops, "sort_stats", {"considered": 0, "reordered": 0})
        torch.manual_seed(11)
        coords, feats = synth.random_cloud(7, 6000, extent=30, n_batch=2, channels=27)
        labels = torch.from_numpy(np.random.default_rng(3).integers(0, 20, size=coords.shape[0]))
        model = models.Res16UNet14A(27, 20).train()
        field = _field(coords, feats)
        out = model(field)
        loss = ops.cross_entropy(out, labels)
        loss.backward()
        mgr = field.coordinate_manager
        return (out.detach().clone(), float(loss), {n: p.grad.clone() for n, p in model.named_parameters()},
                dict(ops.sort_stats), mgr, fake)

    out_a, loss_a, g_a, stats_a, mgr_a, _ = run(False)
    out_b, loss_b, g_b, stats_b, mgr_b, fake_b = run(True)
    assert stats_a["reordered"] == 0 and stats_b["reordered"] >= 1, (stats_a, stats_b)
    assert fake_b.calls.count("spc_table_relabel") == stats_b["reordered"]
    assert torch.allclose(out_a, out_b, rtol=1e-4, atol=1e-5)
    assert abs(loss_a - loss_b) <= 1e-5 * abs(loss_a)
    for name in g_a:
        assert _cos(g_a[name], g_b[name]) >= 0.9999, name
    # the re-ordered map holds the same voxels, and a stride-2 level built from it is consistent with it
    k1 = mgr_b.get_unique_coordinate_map_key(1)
    ca, cb = mgr_a.get_coordinates(mgr_a.get_unique_coordinate_map_key(1)), mgr_b.get_coordinates(k1)
    assert not torch.equal(ca, cb)
    assert sorted(map(tuple, ca.tolist())) == sorted(map(tuple, cb.tolist()))
    k2 = mgr_b.get_unique_coordinate_map_key(2)
    first2, inv2, cnt2 = mgr_b._insert_aux[k2]
    c2 = mgr_b.get_coordinates(k2)
    want = cb.clone()
    want[:, 1:] = torch.div(want[:, 1:], 2, rounding_mode="floor") * 2
    assert torch.equal(c2[inv2.long()], want)                      # parent row -> child row, in both new numberings
    assert torch.equal(want[first2.long()], c2) and int(cnt2.sum()) == cb.shape[0]
