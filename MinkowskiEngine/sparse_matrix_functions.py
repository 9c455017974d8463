"""`spmm` is imported by the reference's weight-sparse conv (modules/sparse_conv.py:11), which
is out of scope; provided through torch.sparse so the import resolves and the call works."""
import torch


def spmm(rows, cols, vals, size, mat, is_sorted=False, cuda_spmm_alg=1):
    idx = torch.stack([rows.long(), cols.long()])
    sp = torch.sparse_coo_tensor(idx, vals, size=tuple(size))
    return torch.sparse.mm(sp, mat)
