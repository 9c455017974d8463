"""ME-compatible layers backed by the sm_100a kernels.

Constructor signatures, attribute names and state-dict keys follow MinkowskiEngine 0.5.x as
the reference uses them (co3d_3d/src/models/mink/modules/common.py:22-180,
modules/sparse_conv.py:267-452 for parameter shapes / init, resnet.py:18,61-64 for pooling).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn
from torch.nn import Parameter

from .. import lib as L
from .. import ops
from .core import (ConvolutionMode, CoordinateMapKey, KernelGenerator, RegionType, SparseTensor, TensorField,
                   _Pending, _to_list)


class MinkowskiModuleBase(nn.Module):
    pass


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


def _wrap_like(input, feats):
    """Re-wrap features with the key/manager of `input` (layernorm.py:16-30 pattern)."""
    if isinstance(input, TensorField):
        return TensorField(feats, coordinate_field_map_key=input.coordinate_field_map_key,
                           coordinate_manager=input.coordinate_manager, quantization_mode=input.quantization_mode)
    return SparseTensor(feats, coordinate_map_key=input.coordinate_map_key,
                        coordinate_manager=input.coordinate_manager)


# ---------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------
class MinkowskiConvolutionBase(MinkowskiModuleBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, is_transpose=False, expand_coordinates=False,
                 convolution_mode=ConvolutionMode.DEFAULT, dimension=-1):
        super().__init__()
        assert dimension > 0, f"Invalid dimension. Please provide a valid dimension argument. dimension={dimension}"
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride, dilation=dilation,
                                               expand_coordinates=expand_coordinates, dimension=dimension)
        elif kernel_generator.expand_coordinates != expand_coordinates:
            kernel_generator.expand_coordinates = expand_coordinates
        if expand_coordinates:
            raise NotImplementedError("expand_coordinates (generative transposed conv) is not built")
        self.is_transpose = is_transpose
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_generator = kernel_generator
        self.dimension = dimension
        self.use_mm = False  # kernel_volume == 1 and all strides 1 -> plain matrix product
        self.convolution_mode = convolution_mode
        # 'tf32' / 'bf16' / 'fp32' / None (= ops.default_precision())
        self.precision: Optional[str] = None
        Tensor = torch.FloatTensor
        if self.kernel_generator.kernel_volume == 1 and self.kernel_generator.requires_strided_coordinates:
            kernel_shape = (self.in_channels, self.out_channels)
            self.use_mm = True
        else:
            kernel_shape = (self.kernel_generator.kernel_volume, self.in_channels, self.out_channels)
        self.kernel = Parameter(Tensor(*kernel_shape))
        self.bias = Parameter(Tensor(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self, is_transpose=False):
        # sparse_conv.py:427-435
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) * self.kernel_generator.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.data.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.data.uniform_(-stdv, stdv)

    def _precision(self) -> int:
        if self.precision is None:
            return ops.default_precision()
        return ops.PRECISIONS[self.precision]

    def forward(self, input: SparseTensor, coordinates=None):
        assert isinstance(input, SparseTensor), "MinkowskiConvolution expects a SparseTensor"
        assert input.D == self.dimension
        if coordinates is not None:
            raise NotImplementedError("explicit output coordinates are not built")
        mgr = input.coordinate_manager
        in_key = input.coordinate_map_key
        kg = self.kernel_generator
        if self.use_mm:
            out_key = in_key
            km = mgr.identity_map(in_key)
            w = self.kernel.view(1, self.in_channels, self.out_channels)
        else:
            if self.is_transpose:
                # existing map at tensor_stride / stride with id "" (sparse_conv.py:397-401)
                ts = input.tensor_stride
                out_ts = []
                for t, s in zip(ts, kg.kernel_stride):
                    if t % s != 0:
                        raise RuntimeError(f"tensor stride {ts} is not divisible by upsample stride {kg.kernel_stride}")
                    out_ts.append(t // s)
                out_key = CoordinateMapKey(out_ts, "")
                if not mgr.exists_coordinate_map_key(out_key):
                    raise RuntimeError(f"transposed convolution needs an existing coordinate map {out_key}")
            else:
                out_key = mgr.stride(in_key, kg.kernel_stride)
            km = mgr.get_kernel_map(in_key, out_key, kg, is_transpose=self.is_transpose)
            w = self.kernel
        precision = self._precision()
        # a bf16 convolution reads the operand copy of its input rows: pending / hollow rows stay without fp32 image
        x = input._operand() if precision == L.PREC_BF16 else input.F
        conv = (x, w, self.bias, km, precision, self.kernel, None if self.use_mm else self._offset_bits(), input)
        if ops.fuse_conv_bn:
            # the rows stay pending: a following MinkowskiBatchNorm takes the convolution into its autograd node
            return SparseTensor._deferred(_Pending(conv=conv), out_key, mgr)
        return SparseTensor(_Pending(conv=conv).run(), coordinate_map_key=out_key, coordinate_manager=mgr)

    def _offset_bits(self) -> Optional[int]:
        """Kernel offsets that take part (None = all); the weight-sparse subclasses restrict them."""
        return None

    def __repr__(self):
        s = f"(in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_generator.kernel_size}, " \
            f"stride={self.kernel_generator.kernel_stride}, dilation={self.kernel_generator.kernel_dilation})"
        return self.__class__.__name__ + s


class MinkowskiConvolution(MinkowskiConvolutionBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=ConvolutionMode.DEFAULT,
                 dimension=None):
        MinkowskiConvolutionBase.__init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias,
                                          kernel_generator, is_transpose=False,
                                          expand_coordinates=expand_coordinates,
                                          convolution_mode=convolution_mode, dimension=dimension)
        self.reset_parameters()


class MinkowskiConvolutionTranspose(MinkowskiConvolutionBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=ConvolutionMode.DEFAULT,
                 dimension=None):
        MinkowskiConvolutionBase.__init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias,
                                          kernel_generator, is_transpose=True,
                                          expand_coordinates=expand_coordinates,
                                          convolution_mode=convolution_mode, dimension=dimension)
        self.reset_parameters(True)


# ---------------------------------------------------------------------------
# weight-sparse inference convolution (the reference's co3d_3d/src/models/mink/modules/sparse_conv.py:267-452)
# ---------------------------------------------------------------------------
class SparseConvMode:
    """sparse_conv.py:19-25"""
    DENSE, SPARSE, ZAXIS, NATIVE, SKIP, SPARSE_DENSE = 0, 1, 2, 3, 4, 5


class _WeightSparseMixin:
    """`sparsify()` (sparse_conv.py:346-379) records `valid_kernel`, the offsets whose pruned kernel W_k still holds
    a non-zero weight (SparseConvMode.ZAXIS: the three offsets along z, `[4, 13, 22]`, whatever the weights); the
    forward pass then runs the ordinary gather-GEMM-scatter kernels with those offsets ONLY: the reference loops over
    `valid_kernel` with one gather + spmm + scatter each (:122-143), here the per-tile offset mask of the tcgen05
    kernels drops the other offsets' pipeline stages (no gather, no MMA) in the single launch.  Inside a kept offset
    the product is dense — unstructured zeros in W_k do not change what a tensor-core tile costs."""

    def _init_sparse(self, sparse_mode):
        self.sparse_mode = int(getattr(sparse_mode, "value", sparse_mode))
        self.valid_kernel = None
        self._flops = 0

    def sparsify(self, layout: str = "strided"):
        if layout not in ("csr", "coo", "strided"):
            raise ValueError(f"unknown layout {layout}")
        if self.use_mm:
            return
        with torch.no_grad():
            nz = (self.kernel != 0).flatten(1).any(dim=1).tolist()   # one host read at conversion time
        self.valid_kernel = [k for k, v in enumerate(nz) if v]
        if self.sparse_mode == SparseConvMode.ZAXIS:
            self.valid_kernel = [4, 13, 22]                          # sparse_conv.py:375-379

    def _offset_bits(self):
        if self.use_mm:
            return None
        assert self.valid_kernel is not None, "call sparsify() first (sparse_conv.py:389)"
        bits = 0
        for k in self.valid_kernel:
            bits |= 1 << k
        return bits


class WeightSparseConvolution(_WeightSparseMixin, MinkowskiConvolution):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=ConvolutionMode.DEFAULT,
                 dimension=None, sparse_mode=1):
        MinkowskiConvolution.__init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias,
                                      kernel_generator, expand_coordinates, convolution_mode, dimension)
        self._init_sparse(sparse_mode)


class WeightSparseConvolutionTranspose(_WeightSparseMixin, MinkowskiConvolutionTranspose):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=ConvolutionMode.DEFAULT,
                 dimension=None, sparse_mode=1):
        MinkowskiConvolutionTranspose.__init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias,
                                               kernel_generator, expand_coordinates, convolution_mode, dimension)
        self._init_sparse(sparse_mode)


# ---------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------
class MinkowskiBatchNorm(nn.Module):
    """`self.bn` is a real nn.BatchNorm1d (state-dict keys bn.weight ... bn.num_batches_tracked,
    fcnn.py:138-140); the arithmetic runs in the fused sm_100a BN kernels."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def _bn_args(self, on_cuda: bool, rows: int, module_training=None):
        """(training, momentum, tracked) of this forward pass; advances num_batches_tracked on the host when the
        statistics kernel cannot (cumulative average, CPU harness, empty input).  `module_training`: bn.training as
        it was when the module was called (deferred rows are produced later)."""
        bn = self.bn
        in_training = bn.training if module_training is None else module_training
        training = in_training or not bn.track_running_stats
        momentum = bn.momentum
        tracked = None
        if in_training and bn.track_running_stats and bn.num_batches_tracked is not None:
            if momentum is None:     # cumulative moving average: the host needs the count
                bn.num_batches_tracked.add_(1)
                momentum = 1.0 / float(bn.num_batches_tracked)
            elif on_cuda and rows >= 1:
                tracked = bn.num_batches_tracked   # incremented by the statistics kernel's last block
            else:
                bn.num_batches_tracked.add_(1)
        return training, (0.0 if momentum is None else momentum), tracked

    def _run(self, feats, relu=False, residual=None, want_fp32=True, module_training=None):
        bn = self.bn
        training, momentum, tracked = self._bn_args(feats.is_cuda, feats.shape[0], module_training)
        return ops.BatchNormFn.apply(feats, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                     momentum, bn.eps, relu, residual, tracked, want_fp32)

    def _run_fused(self, x, w, km, w_param, relu, residual, want_fp32, module_training=None, alias=False):
        """convolution + this BatchNorm (+ residual, ReLU) as one autograd node (ops.ConvBNFn)"""
        bn = self.bn
        training, momentum, tracked = self._bn_args(x.is_cuda, km.m_out, module_training)
        return ops.ConvBNFn.apply(x, w, km, w_param, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                  momentum, bn.eps, relu, residual, tracked, want_fp32, alias)

    def forward(self, input):
        if isinstance(input, SparseTensor):
            # defer the apply pass: a following `+= residual` / ReLU is fused into it (_fused_relu), and a pending
            # convolution that produced `input` joins the node
            pend = input._pending if input._F is None else None
            if pend is not None and pend.bn is None and pend.conv is not None and pend.gone is None:
                mine = _Pending(conv=pend.conv, bn=self)   # (`input` keeps its own pending convolution)
            else:
                mine = _Pending(x=input.F, bn=self)
            return SparseTensor._deferred(mine, input.coordinate_map_key, input.coordinate_manager)
        return _wrap_like(input, self._run(input.F))

    def __repr__(self):
        bn = self.bn
        return f"{self.__class__.__name__}({bn.num_features}, eps={bn.eps}, momentum={bn.momentum}, " \
               f"affine={bn.affine}, track_running_stats={bn.track_running_stats})"


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    """Statistics over the rows of all ranks (train.py:106-107; off by default in the reference,
    `use_sync_batchnorm`).  `self.bn` is an nn.SyncBatchNorm (same state-dict keys, `convert_sync_batchnorm`); in a
    training forward pass with an initialised process group the work is `ops.SyncBatchNormFn` (our statistics / apply
    kernels + one all-reduce of 2C + 1 doubles each way), otherwise it is the ordinary batch norm."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None):
        nn.Module.__init__(self)
        self.bn = nn.SyncBatchNorm(num_features, eps=eps, momentum=momentum, affine=affine,
                                   track_running_stats=track_running_stats, process_group=process_group)

    def forward(self, input):
        import torch.distributed as dist
        bn = self.bn
        synced = bn.training and dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(bn.process_group) > 1
        if not synced:
            return _wrap_like(input, self._run(input.F))
        momentum = bn.momentum
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        out = ops.SyncBatchNormFn.apply(input.F, bn.weight, bn.bias,
                                        bn.running_mean if bn.track_running_stats else None,
                                        bn.running_var if bn.track_running_stats else None,
                                        0.0 if momentum is None else momentum, bn.eps, bn.process_group)
        return _wrap_like(input, out)

    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        module_output = module
        if isinstance(module, MinkowskiBatchNorm) and not isinstance(module, MinkowskiSyncBatchNorm):
            bn = module.bn
            module_output = MinkowskiSyncBatchNorm(bn.num_features, bn.eps, bn.momentum, bn.affine,
                                                   bn.track_running_stats, process_group)
            if bn.affine:
                with torch.no_grad():
                    module_output.bn.weight = bn.weight
                    module_output.bn.bias = bn.bias
            module_output.bn.running_mean = bn.running_mean
            module_output.bn.running_var = bn.running_var
            module_output.bn.num_batches_tracked = bn.num_batches_tracked
            return module_output
        for name, child in module.named_children():
            module_output.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        del module
        return module_output


class MinkowskiInstanceNorm(MinkowskiModuleBase):
    """Per-instance (batch index) normalisation with a [1, C] affine, built by get_norm('IN') (common.py:25-26).
    ME computes it as global-average-pool / broadcast passes: mean over the rows of an instance, biased variance,
    (x - mean) / sqrt(var + 1e-8), then `* weight + bias`; here it is spc_inst_norm_fwd/bwd."""
    EPS = 1e-8

    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features
        self.weight = Parameter(torch.ones(1, num_features))
        self.bias = Parameter(torch.zeros(1, num_features))

    def reset_parameters(self):
        with torch.no_grad():
            self.weight.fill_(1)
            self.bias.zero_()

    def forward(self, input):
        assert isinstance(input, SparseTensor)
        mgr = input.coordinate_manager
        nb = mgr.size(mgr.origin())
        out = ops.InstanceNormFn.apply(input.F, input.C, nb, self.weight, self.bias, self.EPS)
        return SparseTensor(out, coordinate_map_key=input.coordinate_map_key, coordinate_manager=mgr)

    def __repr__(self):
        return f"{self.__class__.__name__}(nchannels={self.num_features})"


# ---------------------------------------------------------------------------
# non-linearities
# ---------------------------------------------------------------------------
class MinkowskiNonlinearityBase(MinkowskiModuleBase):
    MODULE = None

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.module = self.MODULE(*args, **kwargs)

    def forward(self, input):
        return _wrap_like(input, self.module(input.F))

    def __repr__(self):
        return self.__class__.__name__ + "()"


def _fused_relu(input):
    """ReLU; when `input` is a BatchNorm output whose apply pass is still pending, BN (+ residual add)
    + ReLU run as ONE kernel.  The pre-activation rows of `input` are then never produced."""
    pend = input._pending if (isinstance(input, SparseTensor) and input._F is None) else None
    if pend is not None and pend.bn is not None and not pend.relu and pend.gone is None:
        mine = _Pending(conv=pend.conv, x=pend.x, bn=pend.bn)
        mine.res, mine.relu = pend.res, True
        out = SparseTensor._deferred(mine, input.coordinate_map_key, input.coordinate_manager)
        pend.gone = ("the pre-activation features of this BatchNorm output were fused into the "
                     "following ReLU; read .F before applying the ReLU if you need them")
        return out
    return _wrap_like(input, ops.ReLUFn.apply(input.F))


class MinkowskiReLU(MinkowskiNonlinearityBase):
    MODULE = nn.ReLU

    def forward(self, input):
        return _fused_relu(input)


class MinkowskiPReLU(MinkowskiNonlinearityBase):
    MODULE = nn.PReLU


class MinkowskiLeakyReLU(MinkowskiNonlinearityBase):
    MODULE = nn.LeakyReLU


class MinkowskiELU(MinkowskiNonlinearityBase):
    MODULE = nn.ELU


class MinkowskiCELU(MinkowskiNonlinearityBase):
    MODULE = nn.CELU


class MinkowskiSELU(MinkowskiNonlinearityBase):
    MODULE = nn.SELU


class MinkowskiGELU(MinkowskiNonlinearityBase):
    MODULE = nn.GELU


class MinkowskiSigmoid(MinkowskiNonlinearityBase):
    MODULE = nn.Sigmoid


class MinkowskiTanh(MinkowskiNonlinearityBase):
    MODULE = nn.Tanh


class MinkowskiSoftmax(MinkowskiNonlinearityBase):
    MODULE = nn.Softmax


class MinkowskiDropout(MinkowskiNonlinearityBase):
    MODULE = nn.Dropout


class MinkowskiLinear(MinkowskiModuleBase):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        return _wrap_like(input, self.linear(input.F))

    def __repr__(self):
        return f"{self.__class__.__name__}(in_features={self.linear.in_features}, " \
               f"out_features={self.linear.out_features}, bias={self.linear.bias is not None})"


# ---------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------
class MinkowskiLocalPoolingBase(MinkowskiModuleBase):
    AVG = False

    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        assert dimension is not None and dimension > 0, f"invalid dimension: {dimension}"
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride, dilation=dilation,
                                               dimension=dimension)
        self.kernel_generator = kernel_generator
        self.dimension = dimension

    def forward(self, input: SparseTensor, coordinates=None):
        assert isinstance(input, SparseTensor)
        mgr = input.coordinate_manager
        in_key = input.coordinate_map_key
        out_key = mgr.stride(in_key, self.kernel_generator.kernel_stride)
        km = mgr.get_kernel_map(in_key, out_key, self.kernel_generator, is_pool=True)
        out = ops.LocalPoolFn.apply(input.F, km, self.AVG)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)

    def __repr__(self):
        return f"{self.__class__.__name__}(kernel_size={self.kernel_generator.kernel_size}, " \
               f"stride={self.kernel_generator.kernel_stride}, dilation={self.kernel_generator.kernel_dilation})"


class MinkowskiSumPooling(MinkowskiLocalPoolingBase):
    AVG = False


class MinkowskiAvgPooling(MinkowskiLocalPoolingBase):
    AVG = True


class MinkowskiMaxPooling(MinkowskiLocalPoolingBase):
    """Max over the kernel region (the existing neighbours of every output voxel), gradient to the arg-max rows."""

    def forward(self, input: SparseTensor, coordinates=None):
        assert isinstance(input, SparseTensor)
        mgr = input.coordinate_manager
        in_key = input.coordinate_map_key
        out_key = mgr.stride(in_key, self.kernel_generator.kernel_stride)
        km = mgr.get_kernel_map(in_key, out_key, self.kernel_generator, is_pool=True)
        out = ops.LocalMaxPoolFn.apply(input.F, km)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)


def _global_pool_args(input):
    """(manager, origin key, number of instances, int32 coordinates whose column 0 is the instance of each row)."""
    assert isinstance(input, (SparseTensor, TensorField))
    mgr = input.coordinate_manager
    out_key = mgr.origin()
    nb = mgr.size(out_key)
    coords = input.C if isinstance(input, SparseTensor) else mgr.field_batch_coordinates(input.coordinate_field_map_key)
    return mgr, out_key, nb, coords


class MinkowskiGlobalPooling(MinkowskiModuleBase):
    AVG = True

    def __init__(self, mode=None):
        super().__init__()
        self.pooling_mode = mode

    def forward(self, input):
        mgr, out_key, nb, coords = _global_pool_args(input)
        out = ops.GlobalPoolFn.apply(input.F, coords, nb, self.AVG)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class MinkowskiGlobalAvgPooling(MinkowskiGlobalPooling):
    AVG = True


class MinkowskiGlobalSumPooling(MinkowskiGlobalPooling):
    AVG = False


class MinkowskiGlobalMaxPooling(MinkowskiModuleBase):
    def __init__(self, mode=None):
        super().__init__()
        self.pooling_mode = mode

    def forward(self, input):
        mgr, out_key, nb, coords = _global_pool_args(input)
        out = ops.GlobalMaxPoolFn.apply(input.F, coords, nb)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)

    def __repr__(self):
        return self.__class__.__name__ + "()"


class MinkowskiInterpolation(MinkowskiModuleBase):
    """Trilinear sampling of a sparse tensor at continuous coordinates (transforms.py:472,520-528):
    `forward(input, tfield)` with tfield = float [N, 4] (batch, x, y, z) returns the [N, C] features (a plain tensor,
    as ME does); `return_kernel_map` / `return_weights` append the (voxel rows, point rows) pairs and their weights."""

    def __init__(self, return_kernel_map=False, return_weights=False):
        super().__init__()
        self.return_kernel_map = return_kernel_map
        self.return_weights = return_weights

    def forward(self, input: SparseTensor, tfield: torch.Tensor):
        assert isinstance(input, SparseTensor)
        mgr = input.coordinate_manager
        query = tfield.to(device=input.F.device, dtype=torch.float32)
        idx, w = ops.interp_map(mgr._map(input.coordinate_map_key), query)
        out = ops.InterpolateFn.apply(input.F, idx, w)
        if not (self.return_kernel_map or self.return_weights):
            return out
        valid = idx >= 0
        points = torch.arange(idx.shape[1], device=idx.device).expand_as(idx)
        result = [out]
        if self.return_kernel_map:
            result.append((idx[valid].long(), points[valid]))
        if self.return_weights:
            result.append(w[valid])
        return tuple(result)

    def __repr__(self):
        return self.__class__.__name__ + "()"


# ---------------------------------------------------------------------------
# tensor ops
# ---------------------------------------------------------------------------
def cat(*sparse_tensors):
    """Channel concatenation of tensors sharing a coordinate map (res16unet.py:410-425)."""
    if len(sparse_tensors) == 1 and isinstance(sparse_tensors[0], (list, tuple)):
        sparse_tensors = tuple(sparse_tensors[0])
    for s in sparse_tensors:
        assert isinstance(s, (SparseTensor, TensorField)), "Inputs must be sparse tensors."
    first = sparse_tensors[0]
    mgr = first.coordinate_manager
    if isinstance(first, TensorField):
        key = first.coordinate_field_map_key
        for s in sparse_tensors:
            assert mgr is s.coordinate_manager, "different coordinate managers"
            assert key == s.coordinate_field_map_key, "different coordinate field keys"
        return TensorField(torch.cat([s.F for s in sparse_tensors], dim=1), coordinate_field_map_key=key,
                           coordinate_manager=mgr, quantization_mode=first.quantization_mode)
    key = first.coordinate_map_key
    for s in sparse_tensors:
        assert mgr is s.coordinate_manager, \
            "Invalid coordinate manager. All inputs must have the same coordinate manager."
        assert key == s.coordinate_map_key, \
            f"Invalid coordinate map key: {key} != {s.coordinate_map_key}. Inputs must share a coordinate map."
    if ops.lazy_cat and ops.default_precision() == L.PREC_BF16 and all(isinstance(s, SparseTensor) for s in sparse_tensors):
        # the rows stay pending: a bf16 convolution that reads them asks for the bf16 operand copy only (ops.CatFn)
        return SparseTensor._deferred(_Pending(cat=[s._operand() for s in sparse_tensors]), key, mgr)
    return SparseTensor(torch.cat([s.F for s in sparse_tensors], dim=1), coordinate_map_key=key,
                        coordinate_manager=mgr)
