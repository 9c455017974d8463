from nerf_downstream_b200.me.utils import *  # noqa: F401,F403
from nerf_downstream_b200.me.utils import (SparseCollation, batch_sparse_collate, batched_coordinates,  # noqa: F401
                                           kaiming_normal_, sparse_collate, sparse_quantize)
