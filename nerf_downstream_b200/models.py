"""From-scratch definitions of the two networks BASELINE.json names, written against the
ME-compatible surface (`nerf_downstream_b200.me`).

They reproduce the layer tables of the reference models (derived from
co3d_3d/src/models/mink/resnet.py:40-99,107-177 and res16unet.py:72-306,391-435 with
resnet_block.py:14-69) and use the same sub-module names, so a state_dict of the reference's
`ResNet14` / `Res16UNet34C` loads here and vice versa.  The reference files themselves run
unchanged on top of the same surface wherever /root/reference is mounted (tests/test_dropin.py);
these copies exist because the GPU box only receives this repository.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from . import me as ME


def _conv(cin, cout, k, stride=1, bias=False):
    return ME.MinkowskiConvolution(cin, cout, kernel_size=k, stride=stride, dilation=1, bias=bias, dimension=3)


def _conv_tr(cin, cout, k, stride):
    return ME.MinkowskiConvolutionTranspose(cin, cout, kernel_size=k, stride=stride, dilation=1, bias=False,
                                            dimension=3)


def _bn(c, momentum=0.1):
    return ME.MinkowskiBatchNorm(c, momentum=momentum)


class ResidualBlock(nn.Module):
    """conv3-BN-ReLU-conv3-BN (+ shortcut) - add - ReLU   (resnet_block.py:11-69)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride)
        self.norm1 = _bn(planes)
        self.conv2 = _conv(planes, planes, 3, 1)
        self.norm2 = _bn(planes)
        self.downsample = downsample
        self.nonlinearity = ME.MinkowskiReLU()

    def forward(self, x):
        out = self.nonlinearity(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        residual = x if self.downsample is None else self.downsample(x)
        out += residual
        return self.nonlinearity(out)


def _stage(inplanes, planes, n_blocks, stride=1):
    """First block may change stride / width (1x1 conv + BN shortcut), the rest are plain."""
    downsample = None
    if stride != 1 or inplanes != planes:
        downsample = nn.Sequential(_conv(inplanes, planes, 1, stride), _bn(planes))
    blocks = [ResidualBlock(inplanes, planes, stride, downsample)]
    blocks += [ResidualBlock(planes, planes) for _ in range(1, n_blocks)]
    return nn.Sequential(*blocks)


class GlobalAvgPool(nn.Module):
    def __init__(self):
        super().__init__()
        self.global_avg_pool = ME.MinkowskiGlobalAvgPooling()

    def forward(self, x):
        return self.global_avg_pool(x)


class SparseResNet(ME.MinkowskiNetwork):
    """Sparse ResNet classifier: stem conv3 - BN - ReLU - SumPool(2,2) - 4 stride-2 stages -
    global average pool - 1x1 conv with bias.  Returns logits [B, out_channel]."""
    LAYERS: Sequence[int] = (1, 1, 1, 1)
    PLANES: Sequence[int] = (64, 128, 256, 512)
    INIT_DIM = 64

    def __init__(self, in_channel, out_channel, D=3):
        super().__init__(D)
        self.conv1 = _conv(in_channel, self.INIT_DIM, 3, 1)
        self.bn1 = _bn(self.INIT_DIM)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.pool = ME.MinkowskiSumPooling(kernel_size=2, stride=2, dimension=D)
        c = self.INIT_DIM
        for i, (planes, n) in enumerate(zip(self.PLANES, self.LAYERS), start=1):
            setattr(self, f"layer{i}", _stage(c, planes, n, stride=2))
            c = planes
        self.glob_avg = GlobalAvgPool()
        self.final = _conv(c, out_channel, 1, bias=True)

    def process_input(self, batch):
        return ME.TensorField(coordinates=batch["coordinates"], features=batch["features"])

    def forward(self, x):
        out = x.sparse()
        out = self.pool(self.relu(self.bn1(self.conv1(out))))
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        return self.final(self.glob_avg(out)).F


class ResNet14(SparseResNet):
    LAYERS = (1, 1, 1, 1)


class ResNet18(SparseResNet):
    LAYERS = (2, 2, 2, 2)


class ResNet34(SparseResNet):
    LAYERS = (3, 4, 6, 3)


class SparseResUNet(nn.Module):
    """Res16UNet ("MinkUNet"): 2x conv3 stem, 4x [conv2 s2 + residual stage] down, 4x [convtr2 s2 +
    cat(skip) + residual stage] up, 1x1 head, sliced back to the input points."""
    PLANES: Sequence[int] = (32, 64, 128, 256, 256, 128, 96, 96)
    LAYERS: Sequence[int] = (2, 3, 4, 6, 2, 2, 2, 2)

    def __init__(self, in_channel, out_channel, D=3):
        super().__init__()
        P, Ls = self.PLANES, self.LAYERS
        self.D = D
        relu = ME.MinkowskiReLU
        self.conv0p1s1 = nn.Sequential(_conv(in_channel, P[0], 3), _bn(P[0]), relu(),
                                       _conv(P[0], P[0], 3), _bn(P[0]), relu())
        c = P[0]
        enc_names = ["conv1p1s2", "conv2p2s2", "conv3p4s2", "conv4p8s2"]
        skips = [P[0]]
        for i, name in enumerate(enc_names):
            setattr(self, name, nn.Sequential(_conv(c, c, 2, 2), _bn(c), relu()))
            setattr(self, f"block{i + 1}", _stage(c, P[i], Ls[i]))
            c = P[i]
            skips.append(c)
        skips.pop()  # the bottleneck output is not a skip
        dec_names = ["convtr4p16s2", "convtr5p8s2", "convtr6p4s2", "convtr7p2s2"]
        for i, name in enumerate(dec_names):
            planes = P[4 + i]
            setattr(self, name, nn.Sequential(_conv_tr(c, planes, 2, 2), _bn(planes), relu()))
            c = planes + skips.pop()
            setattr(self, f"block{5 + i}", _stage(c, planes, Ls[4 + i]))
            c = planes
        self.final = _conv(c, out_channel, 1, bias=True)

    def forward(self, x: ME.TensorField):
        return self.forward_sparse(x).slice(x).F

    def forward_sparse(self, x: ME.TensorField) -> ME.SparseTensor:
        """The head's voxel logits before `slice` — what `pipeline.seg_head_loss` consumes (one fused pass instead
        of slice -> loss -> metrics)."""
        out = x.sparse()
        skips = [self.conv0p1s1(out)]
        out = skips[0]
        for i, name in enumerate(["conv1p1s2", "conv2p2s2", "conv3p4s2", "conv4p8s2"]):
            out = getattr(self, f"block{i + 1}")(getattr(self, name)(out))
            skips.append(out)
        skips.pop()
        for i, name in enumerate(["convtr4p16s2", "convtr5p8s2", "convtr6p4s2", "convtr7p2s2"]):
            out = ME.cat(getattr(self, name)(out), skips.pop())
            out = getattr(self, f"block{5 + i}")(out)
        return self.final(out)


class Res16UNet34C(SparseResUNet):
    PLANES = (32, 64, 128, 256, 256, 128, 96, 96)
    LAYERS = (2, 3, 4, 6, 2, 2, 2, 2)


class Res16UNet14A(SparseResUNet):
    PLANES = (32, 64, 128, 256, 128, 128, 96, 96)
    LAYERS = (1, 1, 1, 1, 1, 1, 1, 1)


MinkUNet34C = Res16UNet34C


# ---- the other in-tree backbones ------------------------------------------------------------
class GlobalMaxAvgPool(nn.Module):
    """cat(global max, global average) per batch index (fcnn.py:9-18)."""

    def __init__(self):
        super().__init__()
        self.global_max_pool = ME.MinkowskiGlobalMaxPooling()
        self.global_avg_pool = ME.MinkowskiGlobalAvgPooling()

    def forward(self, tensor):
        return ME.cat(self.global_max_pool(tensor), self.global_avg_pool(tensor))


def _mlp_block(cin, cout):
    return nn.Sequential(ME.MinkowskiLinear(cin, cout, bias=False), ME.MinkowskiBatchNorm(cout), ME.MinkowskiLeakyReLU())


def _conv_block(cin, cout, kernel_size, stride, D=3):
    return nn.Sequential(ME.MinkowskiConvolution(cin, cout, kernel_size=kernel_size, stride=stride, dimension=D),
                         ME.MinkowskiBatchNorm(cout), ME.MinkowskiLeakyReLU())


class MinkowskiFCNN(ME.MinkowskiNetwork):
    """Point MLP -> voxel pyramid (conv + max-pool k3 s2, four levels) -> every level sliced back to the points ->
    cat -> three stride-2 convs -> global max+avg -> MLP classifier (fcnn.py:21-168; same sub-module names)."""

    def __init__(self, in_channel, out_channel, kernel_size=3, embedding_channel=1024,
                 channels=(32, 48, 64, 96, 128), D=3):
        super().__init__(D)
        c = channels
        self.mlp1 = _mlp_block(in_channel, c[0])
        self.conv1 = _conv_block(c[0], c[1], kernel_size, 1, D)
        self.conv2 = _conv_block(c[1], c[2], kernel_size, 2, D)
        self.conv3 = _conv_block(c[2], c[3], kernel_size, 2, D)
        self.conv4 = _conv_block(c[3], c[4], kernel_size, 2, D)
        e = embedding_channel
        self.conv5 = nn.Sequential(_conv_block(c[1] + c[2] + c[3] + c[4], e // 4, 3, 2, D),
                                   _conv_block(e // 4, e // 2, 3, 2, D), _conv_block(e // 2, e, 3, 2, D))
        self.max_pool = ME.MinkowskiMaxPooling(kernel_size=3, stride=2, dimension=D)
        self.final = nn.Sequential(GlobalMaxAvgPool(), _mlp_block(e * 2, 512), ME.MinkowskiDropout(),
                                   _mlp_block(512, 512), ME.MinkowskiLinear(512, out_channel, bias=True))
        for m in self.modules():                                    # weight_initialization (fcnn.py:133-140)
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def process_input(self, batch):
        return ME.TensorField(coordinates=batch["coordinates"], features=batch["features"])

    def _to_voxels(self, x):
        return x.sparse()

    def _to_points(self, y, x):
        return y.slice(x)

    def forward(self, x: ME.TensorField):
        x = self.mlp1(x)
        y = self._to_voxels(x)
        y1 = self.max_pool(self.conv1(y))
        y2 = self.max_pool(self.conv2(y1))
        y3 = self.max_pool(self.conv3(y2))
        y4 = self.max_pool(self.conv4(y3))
        x = ME.cat(*[self._to_points(t, x) for t in (y1, y2, y3, y4)])
        return self.final(self.conv5(x.sparse())).F


class MinkowskiSplatFCNN(MinkowskiFCNN):
    """The same network with trilinear splat / interpolate between points and voxels (fcnn.py:170-208)."""

    def _to_voxels(self, x):
        return x.splat()

    def _to_points(self, y, x):
        return y.interpolate(x)


class MinkowskiPointNet(ME.MinkowskiNetwork):
    """PointNet over a coordinate field: five Linear-BN-ReLU blocks on the points, global max per batch index, MLP head
    (pointnet.py:56-109; same sub-module names)."""

    def __init__(self, in_channel, out_channel, embedding_channel=1024, dimension=3):
        super().__init__(dimension)

        def block(cin, cout):
            return nn.Sequential(ME.MinkowskiLinear(cin, cout, bias=False), ME.MinkowskiBatchNorm(cout), ME.MinkowskiReLU())
        self.conv1, self.conv2, self.conv3 = block(in_channel, 64), block(64, 64), block(64, 64)
        self.conv4, self.conv5 = block(64, 128), block(128, embedding_channel)
        self.max_pool = ME.MinkowskiGlobalMaxPooling()
        self.linear1 = block(embedding_channel, 512)
        self.dp1 = ME.MinkowskiDropout()
        self.linear2 = ME.MinkowskiLinear(512, out_channel, bias=True)

    def process_input(self, batch):
        return ME.TensorField(coordinates=batch["coordinates"], features=batch["features"])

    def forward(self, x: ME.TensorField):
        x = self.conv5(self.conv4(self.conv3(self.conv2(self.conv1(x)))))
        x = self.max_pool(x)
        return self.linear2(self.dp1(self.linear1(x))).F


MODELS = {"ResNet14": ResNet14, "ResNet18": ResNet18, "ResNet34": ResNet34, "Res16UNet34C": Res16UNet34C,
          "MinkUNet34C": Res16UNet34C, "Res16UNet14A": Res16UNet14A, "MinkowskiFCNN": MinkowskiFCNN,
          "MinkowskiSplatFCNN": MinkowskiSplatFCNN, "MinkowskiPointNet": MinkowskiPointNet}
