"""Enum surface of ME's native module that the reference imports directly
(co3d_3d/src/models/mink/modules/sparse_conv.py:12)."""
from nerf_downstream_b200.me.core import (ConvolutionMode, CoordinateMapKey, CoordinateMapType,  # noqa: F401
                                          GPUMemoryAllocatorType, MinkowskiAlgorithm, PoolingMode, RegionType)
