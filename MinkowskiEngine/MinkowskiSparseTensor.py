from nerf_downstream_b200.me.core import (CoordinateMapKey, SparseTensor, SparseTensorOperationMode,  # noqa: F401
                                          SparseTensorQuantizationMode, TensorField)
