"""ME-compatible Python surface of the B200 sparse-convolution engine."""
from . import functional as MinkowskiFunctional  # noqa: F401
from . import utils  # noqa: F401
from .core import (ConvolutionMode, CoordinateManager, CoordinateMapKey, CoordinateMapType,  # noqa: F401
                   GPUMemoryAllocatorType, KernelGenerator, MinkowskiAlgorithm, PoolingMode, RegionType,
                   SparseTensor, SparseTensorOperationMode, SparseTensorQuantizationMode, TensorField)
from .modules import (MinkowskiAvgPooling, MinkowskiBatchNorm, MinkowskiCELU, MinkowskiConvolution,  # noqa: F401
                      MinkowskiConvolutionBase, MinkowskiConvolutionTranspose, MinkowskiDropout, MinkowskiELU,
                      MinkowskiGELU, MinkowskiGlobalAvgPooling, MinkowskiGlobalMaxPooling, MinkowskiGlobalPooling,
                      MinkowskiGlobalSumPooling, MinkowskiInstanceNorm, MinkowskiInterpolation, MinkowskiLeakyReLU, MinkowskiLinear,
                      MinkowskiMaxPooling, MinkowskiModuleBase, MinkowskiNetwork, MinkowskiNonlinearityBase,
                      MinkowskiPReLU, MinkowskiReLU, MinkowskiSELU, MinkowskiSigmoid, MinkowskiSoftmax,
                      MinkowskiSumPooling, MinkowskiSyncBatchNorm, MinkowskiTanh, SparseConvMode,
                      WeightSparseConvolution, WeightSparseConvolutionTranspose, cat)

__version__ = "0.5.4+b200"
