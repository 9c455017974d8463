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
