"""CPU oracle of `latticenet_py.lattice.lattice_funcs` (imported with `*` by
seq_lattice/lattice_modules.py:14 and seq_lattice/models.py:6).

TEST INFRASTRUCTURE ONLY (part of oracle/).  PARITY UNPINNED -- see oracle/lattice_oracle.c.
Every op is written as differentiable torch indexing over integer tables produced by the scalar C
oracle, so torch autograd is the backward oracle as well.  `X.apply(...)` mirrors the call shape of
the upstream autograd Functions (lattice_modules.py:301,304).
"""
import torch

FEXT = 9


def im2row_from_table(values, nbr):
    """out[v, s*C + c] = values[nbr[v,s], c]; zeros where nbr is -1 or has no value row yet."""
    V, C = values.shape
    pad = torch.cat([values, values.new_zeros(1, C)], 0)
    idx = torch.where((nbr < 0) | (nbr >= V), torch.full_like(nbr, V), nbr)
    return pad[idx.reshape(-1)].reshape(nbr.shape[0], FEXT * C)


class Im2RowLattice:
    @staticmethod
    def apply(values, ls, filter_extent, dilation, nr_filters):
        assert filter_extent == FEXT
        return im2row_from_table(values, ls.neighbours(dilation=dilation))


class Im2RowIndicesLattice:
    """[V, 9*nr_filters]: neighbour vertex id replicated nr_filters times, -1 = absent; the centre
    slot holds the vertex's own id (convention U5; lattice_modules.py:304,318)."""

    @staticmethod
    def apply(values, ls, filter_extent, dilation, nr_filters):
        nbr = ls.neighbours(dilation=dilation)
        return nbr.to(torch.int32).repeat_interleave(nr_filters, dim=1)


class ConvIm2RowLattice:
    @staticmethod
    def apply(values, ls, weight, dilation):
        return im2row_from_table(values, ls.neighbours(dilation=dilation)).mm(weight)


class CoarsenLattice:
    @staticmethod
    def apply(values_fine, ls_fine, weight):
        coarse = ls_fine.create_coarse_verts()
        nbr = coarse.neighbours(ls_fine, mode=1)
        return im2row_from_table(values_fine, nbr).mm(weight), coarse


class FinefyLattice:
    @staticmethod
    def apply(values_coarse, ls_coarse, ls_fine, weight):
        nbr = ls_fine.neighbours(ls_coarse, mode=2)
        return im2row_from_table(values_coarse, nbr).mm(weight)


def _rows(values, idx):
    V, C = values.shape
    pad = torch.cat([values, values.new_zeros(1, C)], 0)
    i = idx.long()
    i = torch.where((i < 0) | (i >= V), torch.full_like(i, V), i)
    return pad[i], (i < V)


class GatherLattice:
    """[N, 4*(C+1)]: per simplex vertex [w*v(0..C-1), w]; absent vertex -> zeros (convention U6)."""

    @staticmethod
    def apply(values, ls, positions, indices, weights):
        r, ok = _rows(values, indices)
        w = (weights * ok.to(weights.dtype)).unsqueeze(1)
        g = torch.cat([r * w, w], 1)
        return g.reshape(positions.shape[0], -1)


class SliceLattice:
    @staticmethod
    def apply(values, ls, positions, indices, weights):
        r, _ = _rows(values, indices)
        return (r * weights.unsqueeze(1)).reshape(positions.shape[0], 4, -1).sum(1)


class SliceClassifyLattice:
    """logit[p,k] = b[k] + sum_c W[k,c] * sum_r (w+dw)[p,r] * lv[idx[p,r], c]  (SURVEY B.8)."""

    @staticmethod
    def apply(values, ls, positions, delta_weights, linear_weight, linear_bias, nr_classes, indices, weights):
        r, _ = _rows(values, indices)
        ww = weights.reshape(-1, 4) + delta_weights
        s = (r.reshape(positions.shape[0], 4, -1) * ww.unsqueeze(2)).sum(1)
        return s.mm(linear_weight.t()) + linear_bias


class SplatLattice:
    """values[idx] += w * [val, 1]  (homogeneous coordinate last)."""

    @staticmethod
    def apply(ls, positions, values):
        rows, idx, w = ls.distribute(positions, torch.zeros(positions.shape[0], 1), True)
        V = ls.nr_lattice_vertices()
        i = idx.long()
        ok = i >= 0
        src = torch.cat([values, torch.ones(values.shape[0], 1)], 1).repeat_interleave(4, 0) * w.unsqueeze(1)
        out = torch.zeros(V, values.shape[1] + 1).index_add(0, i[ok], src[ok])
        return out, idx, w


class DistributeLattice:
    @staticmethod
    def apply(ls, positions, values, reset_hashmap=True):
        return ls.distribute(positions, values, reset_hashmap)
