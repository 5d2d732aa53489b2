"""CPU restatement of the three torch_scatter (tag 2.0.4, seq_docker/Dockerfile:85-90) functions the
reference calls (seq_lattice/lattice_modules.py:485-520, seq_lattice/models.py:454).

TEST INFRASTRUCTURE ONLY (part of oracle/).  Semantics restated from torch_scatter 2.0.4:
  * output length along `dim` is max(index)+1 unless `out`/`dim_size` is given (quirk Q8);
  * scatter_max: empty segments give out = 0 and argmax = src.size(dim) (out-of-range sentinel, Q3);
  * ties: the smallest source row wins (upstream is a race; fixed here and in the CUDA path);
  * gradients flow to the arg-max rows only.
"""
import torch


def _dim_size(index, dim_size, out, dim):
    if out is not None:
        return out.shape[dim]
    if dim_size is not None:
        return int(dim_size)
    return int(index.max().item()) + 1 if index.numel() > 0 else 0


def _expand(index, src, dim):
    if index.dim() == src.dim():
        return index
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    n = _dim_size(index, dim_size, out, dim)
    idx = _expand(index, src, dim)
    if out is None:
        shape = list(src.shape)
        shape[dim] = n
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.scatter_add(dim, idx, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    n = _dim_size(index, dim_size, out, dim)
    total = scatter_add(src, index, dim, out, n)
    ones = torch.ones(index.shape, dtype=src.dtype, device=src.device)
    count = torch.zeros(n, dtype=src.dtype, device=src.device).scatter_add(0, index, ones).clamp(min=1)
    shape = [1] * src.dim()
    shape[dim] = -1
    return total / count.view(shape)


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    assert dim == 0 and src.dim() == 2, "oracle restates the [R,C], dim=0 case the reference uses"
    n = _dim_size(index, dim_size, out, dim)
    R, C = src.shape
    idx = _expand(index, src, dim)
    with torch.no_grad():
        best = torch.full((n, C), float("-inf"), dtype=src.dtype)
        best = best.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
        rows = torch.arange(R, dtype=torch.long).view(-1, 1).expand(R, C)
        hit = src.detach() == best.gather(0, idx)
        cand = torch.where(hit, rows, torch.full_like(rows, R))
        arg = torch.full((n, C), R, dtype=torch.long).scatter_reduce(0, idx, cand, reduce="amin", include_self=True)
        empty = arg >= R
    val = src.gather(0, arg.clamp(max=R - 1))
    val = torch.where(empty, torch.zeros_like(val), val)
    return val, arg
