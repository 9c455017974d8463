"""The reference's own model files import and construct UNCHANGED on the ME-compatible surface,
and their parameters line up one to one with this repository's from-scratch definitions (so the GPU
parity tests on those definitions cover the reference models).  Skips where /root/reference is not
mounted (the GPU box)."""
import pytest
import torch

from tests import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted")


def _shapes(model):
    return {k: tuple(v.shape) for k, v in model.state_dict().items()}


def test_reference_resnet14_constructs_and_matches():
    ref = ref_harness.load("co3d_3d.src.models.mink.resnet")
    from nerf_downstream_b200 import models
    theirs = ref.ResNet14(in_channel=27, out_channel=51)
    ours = models.ResNet14(27, 51)
    assert _shapes(theirs) == _shapes(ours)
    ours.load_state_dict(theirs.state_dict())
    import MinkowskiEngine as ME
    assert isinstance(theirs.conv1, ME.MinkowskiConvolution) and isinstance(theirs.pool, ME.MinkowskiSumPooling)


def test_reference_res16unet34c_constructs_and_matches():
    ref = ref_harness.load("co3d_3d.src.models.mink.res16unet")
    from nerf_downstream_b200 import models
    theirs = ref.Res16UNet34C(in_channel=27, out_channel=20)
    ours = models.Res16UNet34C(27, 20)
    assert _shapes(theirs) == _shapes(ours)
    theirs.load_state_dict(ours.state_dict())


def test_reference_resunet2_and_sparse_conv_module_import():
    # resunet.py is the only in-tree user of 3^3 stride-2 down / up convolutions
    ref = ref_harness.load("co3d_3d.src.models.mink.resunet")
    assert hasattr(ref, "ResUNet2")
    sc = ref_harness.load("co3d_3d.src.models.mink.modules.sparse_conv")
    assert hasattr(sc, "WeightSparseConvolutionFunction")


def test_reference_fcnn_and_pointnet_construct_and_match():
    """The other in-tree backbones construct unchanged on the surface — they need max pooling,
    global max pooling, Linear / LeakyReLU / Dropout wrappers, splat / interpolate — and their state dicts line up
    with the from-scratch mirrors the GPU tests run."""
    from nerf_downstream_b200 import models
    fc = ref_harness.load("co3d_3d.src.models.mink.fcnn")
    pn = ref_harness.load("co3d_3d.src.models.mink.pointnet")
    for theirs, ours in [(fc.MinkowskiFCNN(27, 51), models.MinkowskiFCNN(27, 51)),
                         (fc.MinkowskiSplatFCNN(27, 51), models.MinkowskiSplatFCNN(27, 51)),
                         (fc.MinkowskiFCNN(3, 40, embedding_channel=64, channels=(8, 16, 16, 32, 32)),
                          models.MinkowskiFCNN(3, 40, embedding_channel=64, channels=(8, 16, 16, 32, 32))),
                         (pn.MinkowskiPointNet(27, 51), models.MinkowskiPointNet(27, 51))]:
        assert _shapes(theirs) == _shapes(ours), type(theirs).__name__
        ours.load_state_dict(theirs.state_dict())
        theirs.load_state_dict(ours.state_dict())


def test_reference_resunet_variants_construct():
    """ResUNetBN2* / ResUNetIN2* (resunet.py:244-300): batch-norm and instance-norm variants, 3^3 stride-2 down / up."""
    import MinkowskiEngine as ME
    ref = ref_harness.load("co3d_3d.src.models.mink.resunet")
    for name in ("ResUNetBN2", "ResUNetBN2C", "ResUNetIN2", "ResUNetIN2C", "ResUNetIN2E"):
        net = getattr(ref, name)(in_channel=27, out_channel=20)
        block_norm = ME.MinkowskiInstanceNorm if "IN" in name else ME.MinkowskiBatchNorm     # BLOCK_NORM_TYPE
        assert isinstance(net.norm1, ME.MinkowskiBatchNorm) and isinstance(net.block1.norm1, block_norm)
        assert isinstance(net.conv4_tr, ME.MinkowskiConvolutionTranspose)
        assert sum(p.numel() for p in net.parameters()) > 7_000_000
    with pytest.raises(ValueError, match="not supported"):
        ref.ResUNet2(in_channel=27, out_channel=20)            # NORM_TYPE None (common.py:22-33), as with ME itself
