"""Autograd Functions over the C ABI -- the `latticenet_py.lattice.lattice_funcs` surface
(/root/reference/seq_lattice/lattice_modules.py:14,301,304; models.py:6).

Forward and backward both run hand-written sm_100a kernels; dense contractions go through
`ops.matmul` (fp32).  No CPU fallback.
"""
import os

import torch
from torch.autograd import Function

from . import _lib
from . import ops

FEXT = 9


def _f32c(t):
    return t.contiguous().float()


# ---------------------------------------------------------------------------------------------
# raw kernels
# ---------------------------------------------------------------------------------------------
def im2row_raw(values, nbr, nr_rows=None):
    values = _f32c(values)
    vq = nbr.shape[0] if nr_rows is None else nr_rows
    C = values.shape[1]
    out = torch.empty(vq, FEXT * C, dtype=torch.float32, device=values.device)
    p = _lib.ptr
    _lib.check(_lib.load().ltn_im2row(p(values), values.shape[0], None, p(nbr), vq, None, C, p(out), _lib.stream()),
               "ltn_im2row")
    return out


def row2im_raw(grad_rows, nbr_t, nr_vals, C):
    grad_rows = _f32c(grad_rows)
    out = torch.zeros(nr_vals, C, dtype=torch.float32, device=grad_rows.device)
    vu = min(nr_vals, nbr_t.shape[0])
    p = _lib.ptr
    _lib.check(_lib.load().ltn_row2im(p(grad_rows), grad_rows.shape[0], p(nbr_t), vu, C, p(out), _lib.stream()),
               "ltn_row2im")
    return out


_FLIP_DEV = {}
_FLIP = (1, 0, 3, 2, 5, 4, 7, 6, 8)   # slot pairs swap under transposition (2a <-> 2a+1), the centre stays


def transposed_weight(weight, C, F):
    """[9C, F] slot-major conv weight -> the weight of the TRANSPOSED convolution, [9F, C]:
    Wt[s'*F + f, c] = W[flip(s')*C + c, f]; cached per parameter version (ops.k_major on the result)."""
    def make():
        # the slot permutation as a cached DEVICE index: indexing with a Python list builds and uploads an index tensor per
        # call (0.2 ms of host time each, 60 convolutions per training step: 13 ms of a 59 ms step in the host profile)
        flip = _FLIP_DEV.get(weight.device)
        if flip is None:
            flip = _FLIP_DEV[weight.device] = torch.tensor(_FLIP, dtype=torch.long, device=weight.device)
        wt = weight.detach().view(FEXT, C, F).index_select(0, flip).transpose(1, 2).contiguous().view(FEXT * F, C)
        return ops.SplitWeight(wt.float(), False)
    return ops._cached_split(weight, (id(weight), "transposed"), make)


class _GatherConv(Function):
    """out = im2row(values, nbr) @ weight.  nbr_t is the opposite-direction table used to run the
    transpose as a gather (see csrc/ltn_gather.cu::k_row2im).  The [V,9C] buffer is not kept for
    backward; it is rebuilt (bandwidth is cheaper than holding 200 MB per conv alive).

    Backward wrt the values (`ltn_conv_bwd_data` of SURVEY.md 8b) is the FORWARD tensor-core kernel run over the
    transposed neighbour table with the transposed weight: grad_x[u, c] = sum_s' sum_f grad_out[nbr_t[u, s'], f] *
    W[flip(s')*C + c, f] -- no [V, 9C] gradient buffer, no row2im pass.  Shapes the tensor-core kernel does not take
    (F % 32 != 0) keep the materialised path."""

    @staticmethod
    def forward(ctx, values, weight, nbr, nbr_t):
        rows = im2row_raw(values, nbr)
        out = ops.matmul(rows, weight)
        ctx.save_for_backward(values, weight, nbr, nbr_t)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        values, weight, nbr, nbr_t = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        gv = gw = None
        C, F = values.shape[1], weight.shape[1]
        if ctx.needs_input_grad[0]:
            if ops.conv_tc_supported(F, C, False) and grad_out.shape[0] > 0 and values.shape[0] > 0:
                vu = min(values.shape[0], nbr_t.shape[0])
                gv = torch.zeros(values.shape[0], C, dtype=torch.float32, device=grad_out.device) if vu < values.shape[0] else \
                    torch.empty(values.shape[0], C, dtype=torch.float32, device=grad_out.device)
                ops.conv_tc(grad_out, nbr_t, transposed_weight(weight, C, F), nr_rows=vu, out=gv, operands="tf32")
            else:
                grad_rows = ops.matmul(grad_out, weight.t())
                gv = row2im_raw(grad_rows, nbr_t, values.shape[0], values.shape[1])
        if ctx.needs_input_grad[1]:
            rows = im2row_raw(values, nbr)
            gw = ops.matmul(rows.t(), grad_out)
        return gv, gw, None, None


def conv_bwd_weight(act, gy, nbr):
    """[S*C, F] weight gradient on the tensor cores (csrc/ltn_conv_bwd_weight.cu): sum_v gathered-act[v]^T gy[v]; None when the
    shape is not supported (the caller falls back to im2row + fp32 GEMM).  nbr None: S = 1."""
    C, F = act.shape[1], gy.shape[1]
    if (C % 4 or F % 16 or F < 16 or F > 256 or gy.shape[0] == 0 or act.shape[0] == 0
            or os.environ.get("LTN_TRAIN_DW_SIMT", "0") == "1"):
        return None
    S = 1 if nbr is None else nbr.shape[1]
    act, gy = _f32c(act), _f32c(gy)
    out = torch.zeros(S * C, F, dtype=torch.float32, device=act.device)
    p = _lib.ptr
    _lib.check(_lib.load().ltn_conv_bwd_weight(p(act), act.shape[0], p(nbr), gy.shape[0], C, S, p(gy), F, p(out), _lib.stream()),
               "ltn_conv_bwd_weight")
    return out


class _FusedConv(Function):
    """Training-time form of a whole layer  y = conv(act(x)) (+ bias) (+ res),  act = relu(GroupNorm(x)) or the identity
    (GnReluConv / GnReluCoarsen / GnReluFinefy / GnRelu1x1 / ConvLatticeModule, lattice_modules.py:75-140,436-440,573).

    forward  = the SAME fused tensor-core kernel inference uses (GroupNorm folded into the gathered operand, bias, residual
               and the output's GroupNorm statistics in the epilogue; tf32 hi/lo operands, three passes, fp32 results): no
               [V, 9C] buffer, no separate normalisation pass.  Round 1 ran im2row + an fp32 SIMT GEMM + two GroupNorm
               kernels here.
    backward = act recomputed in one pass (k_gn_apply) instead of kept alive; d act = the forward kernel over the
               opposite-direction neighbour table with the transposed weight (`ltn_conv_bwd_data` of SURVEY.md 8b);
               d W = gathered-act^T . dy; then the GroupNorm + ReLU backward kernel (k_gn_bwd) gives d x, d gamma, d beta.
    `linear`: weight is an nn.Linear weight [F, C] and there is no neighbour table (S = 1)."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, gamma, beta, nbr, nbr_t, cfg):
        linear, groups, eps, sums_in = cfg
        x = _f32c(x)
        has_gn = gamma is not None
        F = weight.shape[0] if linear else weight.shape[1]
        sums = None
        if has_gn:
            sums = sums_in if sums_in is not None else ops.gn_sums(x, groups)
            if sums._base is not None:
                sums = sums.clone()   # a slot of the per-frame arena: the next frame reuses it, backward comes after all frames
        wt = ops.k_major(weight, transposed=True) if linear else ops.k_major(weight)
        # own storage, not the per-frame arena: back-propagation through time reads these after the later frames have run
        out_sums = torch.zeros(ops.gn_groups(F), 2, dtype=torch.float64, device=x.device)
        out = ops.conv_tc(x, nbr, wt, gn=(sums, gamma.detach(), beta.detach(), eps) if has_gn else None, relu=has_gn,
                          bias=None if bias is None else bias.detach(), res=None if res is None else _f32c(res.detach()),
                          out_sums=out_sums, operands="tf32")
        ctx.save_for_backward(x, weight, gamma, beta, nbr, nbr_t, sums)
        ctx.cfg = (linear, groups, eps, has_gn)
        ctx.mark_non_differentiable(out_sums)
        return out, out_sums

    @staticmethod
    def backward(ctx, gy, _unused):
        x, weight, gamma, beta, nbr, nbr_t, sums = ctx.saved_tensors
        linear, groups, eps, has_gn = ctx.cfg
        need = ctx.needs_input_grad
        gy = _f32c(gy)
        V, C = x.shape
        F = weight.shape[0] if linear else weight.shape[1]
        lib, p = _lib.load(), _lib.ptr
        g_x = g_w = g_b = g_res = g_gamma = g_beta = None
        if need[3]:
            g_res = gy
        if need[2]:
            g_b = gy.sum(0)
        act = x
        if has_gn:
            act = torch.empty_like(x)
            _lib.check(lib.ltn_gn_apply(p(x), V, None, C, groups, p(sums), p(gamma), p(beta), float(eps), 1, p(act), _lib.stream()),
                       "ltn_gn_apply")
        if need[1]:
            g_w = conv_bwd_weight(act, gy, None if linear else nbr)
            if g_w is None:   # shapes the tensor-core kernel does not take
                g_w = ops.matmul(gy.t(), act) if linear else ops.matmul(im2row_raw(act, nbr).t(), gy)
            elif linear:
                g_w = g_w.t().contiguous()                         # nn.Linear layout [F, C]
        if need[0] or (has_gn and (need[4] or need[5])):
            if linear:
                if ops.conv_tc_supported(F, C, False) and os.environ.get("LTN_TRAIN_LIN_DACT_SIMT", "0") != "1":
                    wt_t = ops._cached_split(weight, (id(weight), "linear_t"),
                                             lambda: ops.SplitWeight(weight.detach().t().contiguous().float(), True))
                    d_act = ops.conv_tc(gy, None, wt_t, operands="tf32")
                else:
                    d_act = ops.matmul(gy, weight)
            elif ops.conv_tc_supported(F, C, False) and gy.shape[0] > 0:
                vu = min(V, nbr_t.shape[0])
                d_act = torch.zeros(V, C, dtype=torch.float32, device=gy.device) if vu < V else \
                    torch.empty(V, C, dtype=torch.float32, device=gy.device)
                ops.conv_tc(gy, nbr_t, transposed_weight(weight, C, F), nr_rows=vu, out=d_act, operands="tf32")
            else:
                d_act = row2im_raw(ops.matmul(gy, weight.t()), nbr_t, V, C)
            if has_gn:
                chan = torch.empty(C, 2, dtype=torch.float64, device=x.device)
                g_x = torch.empty_like(x)
                _lib.check(lib.ltn_gn_bwd(p(x), p(d_act), p(act), V, C, groups, p(sums), p(gamma), float(eps), p(chan), p(g_x),
                                          _lib.stream()), "ltn_gn_bwd")
                chan = chan.float()
                g_gamma, g_beta = chan[:, 1].contiguous(), chan[:, 0].contiguous()
            else:
                g_x = d_act
        return g_x, g_w, g_b, g_res, g_gamma, g_beta, None, None, None


def fused_conv_train(x, weight, bias, res, norm, nbr, nbr_t_fn, linear=False):
    """the layer through _FusedConv; `norm` = GroupNormLatticeModule (GroupNorm + ReLU folded in) or None; the output's
    GroupNorm statistics travel with it to the next layer (ops.sums_of), as on the inference path"""
    gamma = beta = sums = None
    groups, eps = 0, 0.0
    if norm is not None:
        gamma, beta, groups, eps = norm.gn.weight, norm.gn.bias, norm.groups, norm.gn.eps
        s = getattr(x, "_ltn_gn_sums", None)
        if s is not None and s[0] == ops._FRAME["id"] and s[1].shape[0] == groups and os.environ.get("LTN_TRAIN_NOSUMS", "0") != "1":
            sums = s[1]
    nbr_t = None
    if not linear:
        # the opposite-direction table is what d act runs over: needed for d x and, through the GroupNorm backward, for d gamma / d beta
        nbr_t = nbr_t_fn() if (x.requires_grad or norm is not None) else nbr
    out, out_sums = _FusedConv.apply(x, weight, bias, res, gamma, beta, nbr, nbr_t, (linear, groups, eps, sums))
    out._ltn_gn_sums = (ops._FRAME["id"], out_sums)
    return out


def train_fusable(x, C, F, norm=None):
    """grad mode, shapes the tensor-core kernel takes, GroupNorm with affine parameters (or none at all)"""
    return (torch.is_grad_enabled() and x.is_cuda and x.shape[0] > 0 and ops.conv_tc_supported(C, F, norm is not None)
            and (norm is None or norm.gn.weight is not None) and ops._FRAME["arena"] is not None
            and os.environ.get("LTN_TRAIN_UNFUSED", "0") != "1")


def gather_conv(values, weight, nbr, nbr_t_fn):
    """nbr_t_fn() builds the opposite-direction table; it is only evaluated when a gradient with
    respect to `values` can be asked for."""
    needs = torch.is_grad_enabled() and values.requires_grad
    nbr_t = nbr_t_fn() if needs else nbr
    return _GatherConv.apply(values, weight, nbr, nbr_t)


# ---------------------------------------------------------------------------------------------
# reference-named Functions
# ---------------------------------------------------------------------------------------------
class Im2RowLattice(Function):
    """[V, 9*C] neighbour rows, slot 8 = centre, zeros where absent (lattice_modules.py:301,316)."""

    @staticmethod
    def forward(ctx, values, ls, filter_extent, dilation, nr_filters):
        if filter_extent != FEXT:
            raise RuntimeError("filter_extent must be 9")
        nbr = ls.neighbours(dilation=dilation)
        ctx.save_for_backward(nbr)
        ctx.shape = tuple(values.shape)
        return im2row_raw(values, nbr)

    @staticmethod
    def backward(ctx, grad_rows):
        (nbr,) = ctx.saved_tensors
        return row2im_raw(grad_rows, nbr, ctx.shape[0], ctx.shape[1]), None, None, None, None


class Im2RowIndicesLattice(Function):
    """[V, 9*nr_filters] int32: neighbour id replicated nr_filters times (the reference strides it
    with [:, ::nr_filters], lattice_modules.py:318,325,339); -1 absent; centre = own id (U5)."""

    @staticmethod
    def forward(ctx, values, ls, filter_extent, dilation, nr_filters):
        nbr = ls.neighbours(dilation=dilation)
        return nbr.repeat_interleave(int(nr_filters), dim=1)

    @staticmethod
    def backward(ctx, g):
        return None, None, None, None, None


class ConvIm2RowLattice:
    @staticmethod
    def apply(values, ls, weight, dilation):
        nbr = ls.neighbours(dilation=dilation)
        return gather_conv(values, weight, nbr, lambda: nbr)


class CoarsenLattice:
    @staticmethod
    def apply(values_fine, ls_fine, weight):
        coarse = ls_fine.create_coarse_verts()
        nbr = coarse.neighbours(ls_fine, mode=1)
        return gather_conv(values_fine, weight, nbr, lambda: ls_fine.neighbours(coarse, mode=2)), coarse


class FinefyLattice:
    @staticmethod
    def apply(values_coarse, ls_coarse, ls_fine, weight):
        nbr = ls_fine.neighbours(ls_coarse, mode=2)
        return gather_conv(values_coarse, weight, nbr, lambda: ls_coarse.neighbours(ls_fine, mode=1))


class DistributeLattice:
    @staticmethod
    def apply(ls, positions, values, reset_hashmap=True, subtract_mean=True):
        with torch.no_grad():
            return ls.distribute(positions, values, reset_hashmap, subtract_mean)


class GatherLattice(Function):
    """[N, 4*(C+1)]: per simplex vertex [w*v, w] (convention U6)."""

    @staticmethod
    def forward(ctx, values, ls, positions, indices, weights):
        values = _f32c(values)
        n, C = positions.shape[0], values.shape[1]
        out = torch.empty(n, 4 * (C + 1), dtype=torch.float32, device=values.device)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_gather(p(values), values.shape[0], _lib.rows_dev(values.shape[0]), C, p(indices), p(weights), n,
                                          _lib.rows_dev(n), p(out), _lib.stream()),
                   "ltn_gather")
        ctx.save_for_backward(indices, weights)
        ctx.shape = tuple(values.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        indices, weights = ctx.saved_tensors
        V, C = ctx.shape
        g = _f32c(g)
        gv = torch.zeros(V, C, dtype=torch.float32, device=g.device)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_gather_bwd(p(g), g.shape[0], C, p(indices), p(weights), p(gv), V, _lib.stream()),
                   "ltn_gather_bwd")
        return gv, None, None, None, None


class SliceLattice(Function):
    @staticmethod
    def forward(ctx, values, ls, positions, indices, weights):
        values = _f32c(values)
        n, C = positions.shape[0], values.shape[1]
        out = torch.empty(n, C, dtype=torch.float32, device=values.device)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_slice(p(values), values.shape[0], C, p(indices), p(weights), n, p(out), _lib.stream()),
                   "ltn_slice")
        ctx.save_for_backward(indices, weights)
        ctx.shape = tuple(values.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        indices, weights = ctx.saved_tensors
        V, C = ctx.shape
        g = _f32c(g)
        gv = torch.zeros(V, C, dtype=torch.float32, device=g.device)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_slice_bwd(p(g), g.shape[0], C, p(indices), p(weights), p(gv), V, _lib.stream()),
                   "ltn_slice_bwd")
        return gv, None, None, None, None


class SliceClassifyLattice(Function):
    """logit[p,k] = b[k] + sum_c W[k,c] * sum_r (w+dw)[p,r] * lv[idx[p,r], c]  (models.py:465)."""

    @staticmethod
    def forward(ctx, values, ls, positions, delta_weights, linear_weight, linear_bias, nr_classes, indices, weights):
        values, dw = _f32c(values), _f32c(delta_weights)
        W, b = _f32c(linear_weight), _f32c(linear_bias)
        n, C, K = positions.shape[0], values.shape[1], W.shape[0]
        out = torch.empty(n, K, dtype=torch.float32, device=values.device)
        need_grad = any(ctx.needs_input_grad)
        sliced = torch.empty(n, C, dtype=torch.float32, device=values.device) if need_grad else None
        p = _lib.ptr
        _lib.check(_lib.load().ltn_slice_classify(p(values), values.shape[0], _lib.rows_dev(values.shape[0]), C, p(indices),
                                                  p(weights), p(dw), n, _lib.rows_dev(n), p(W), p(b), K, p(out), p(sliced),
                                                  _lib.stream()), "ltn_slice_classify")
        if need_grad:
            ctx.save_for_backward(values, dw, W, indices, weights, sliced)
        return out

    @staticmethod
    def backward(ctx, g):
        values, dw, W, indices, weights, sliced = ctx.saved_tensors
        g = _f32c(g)
        V, C = values.shape
        n, K = g.shape
        gv = torch.zeros(V, C, dtype=torch.float32, device=g.device)
        gdw = torch.empty(n, 4, dtype=torch.float32, device=g.device)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_slice_classify_bwd(p(g), p(values), V, C, p(indices), p(weights), p(dw), n, p(W), K,
                                                      p(gv), p(gdw), _lib.stream()), "ltn_slice_classify_bwd")
        gW = ops.matmul(g.t(), sliced)
        gb = g.sum(0)
        return gv, None, None, gdw, gW, gb, None, None, None


class SplatLattice:
    """values[idx] += w*[val, 1] (homogeneous coordinate last); returns (lv [V,C+1], idx, w)."""

    @staticmethod
    def apply(ls, positions, values):
        with torch.no_grad():
            values = _f32c(values)
            zeros = torch.zeros(positions.shape[0], 1, dtype=torch.float32, device=positions.device)
            _, idx, w = ls.distribute(positions, zeros, True, False)
            V = ls.nr_lattice_vertices()
            out = torch.zeros(V, values.shape[1] + 1, dtype=torch.float32, device=values.device)
            p = _lib.ptr
            _lib.check(_lib.load().ltn_splat(p(values), values.shape[0], values.shape[1], p(idx), p(w), p(out), V,
                                             _lib.stream()), "ltn_splat")
            return out, idx, w
