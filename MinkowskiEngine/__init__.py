"""Drop-in `MinkowskiEngine` package name for the reference's unchanged model files
(`import MinkowskiEngine as ME`, co3d_3d/src/models/mink/resnet.py:6).  Everything is
re-exported from `nerf_downstream_b200.me`; the arithmetic runs in libsparseconv_b200.so."""
from nerf_downstream_b200.me import *  # noqa: F401,F403
from nerf_downstream_b200.me import __version__, utils  # noqa: F401

from . import MinkowskiCommon  # noqa: F401
from . import MinkowskiCoordinateManager  # noqa: F401
from . import MinkowskiFunctional  # noqa: F401
from . import MinkowskiKernelGenerator  # noqa: F401
from . import MinkowskiOps  # noqa: F401
from . import MinkowskiSparseTensor  # noqa: F401
from . import sparse_matrix_functions  # noqa: F401
