"""`torch_scatter` (tag 2.0.4) entry points the reference calls (seq_lattice/lattice_modules.py:485-520,
models.py:454) on the sm_100a segmented-reduction kernels.  Only the [R,C]/[R], dim=0 forms the
reference uses are supported."""
import torch

from temporal_latticenet_b200 import ops as _ops


def _check(src, index, dim):
    dim = dim % src.dim()
    if dim != 0 or index.dim() != 1 or index.shape[0] != src.shape[0]:
        raise RuntimeError("only scatter over dim 0 with a 1-D index is supported")


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    _check(src, index, dim)
    if out is not None:
        raise RuntimeError("scatter_max(out=...) is not supported")
    return _ops.scatter_max(src, index, dim_size)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    _check(src, index, dim)
    return _ops.scatter_add(src, index, dim_size, out)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    _check(src, index, dim)
    return _ops.scatter_mean(src, index, dim_size, out)
