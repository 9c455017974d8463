"""MinkowskiEngine.MinkowskiFunctional equivalents (common.py:54-70, resunet.py:184-232)."""
import torch.nn.functional as F

from .. import ops
from .modules import _fused_relu, _wrap_like


def _wrap_fn(fn):
    def wrapped(input, *args, **kwargs):
        return _wrap_like(input, fn(input.F, *args, **kwargs))
    wrapped.__name__ = fn.__name__
    return wrapped


def relu(input, *args, **kwargs):
    """ReLU on the feature rows through the sm_100a elementwise kernel."""
    return _fused_relu(input)


leaky_relu = _wrap_fn(F.leaky_relu)
prelu = _wrap_fn(F.prelu)
elu = _wrap_fn(F.elu)
celu = _wrap_fn(F.celu)
selu = _wrap_fn(F.selu)
gelu = _wrap_fn(F.gelu)
sigmoid = _wrap_fn(F.sigmoid)
tanh = _wrap_fn(F.tanh)
softmax = _wrap_fn(F.softmax)
log_softmax = _wrap_fn(F.log_softmax)
dropout = _wrap_fn(F.dropout)
normalize = _wrap_fn(F.normalize)
