"""nn.Modules over the sm_100a kernels -- the `latticenet_py.lattice.lattice_modules` surface the
reference imports with `*` (/root/reference/seq_lattice/lattice_modules.py:15, models.py:7).

Same names, same `(lv, ls)` protocol, same lazy parameter creation and parameter names as the
recipes in SURVEY.md section 2.2 (E4-E13), so a state-dict moves between the reference-driven
oracle run and this implementation unchanged.  GroupNorm+ReLU is one fused op; every convolution
on a level reuses that level's neighbour table.
"""
import math

import torch

from . import _lib
from . import funcs as F_
from . import ops
from .funcs import *  # noqa: F401,F403  (the reference expects the Functions in this namespace too)

_NO_LOCAL_MEAN = ("pointnet_no_local_mean", "pointnet_no_elevate_no_local_mean", "splat")


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _conv_weight(rows, nr_filters):
    """uniform(+-sqrt(3)*gain/sqrt(fan_out)); fan_out of a [rows, filters] tensor is `rows`
    (the rule the reference keeps a copy of at lattice_modules.py:264-272)."""
    bound = math.sqrt(3.0) * math.sqrt(2.0) / math.sqrt(rows)
    return torch.nn.Parameter(torch.empty(rows, nr_filters, device=_dev()).uniform_(-bound, bound))


def _linear(nr_in, nr_out, bias):
    lin = torch.nn.Linear(nr_in, nr_out, bias=bias).to(_dev())
    with torch.no_grad():
        torch.nn.init.kaiming_normal_(lin.weight, mode="fan_in", nonlinearity="relu")
    return lin


class DistributeLatticeModule(torch.nn.Module):
    """models.py:62,297-298: (ls, positions[N,3], values[N,vd], reset) ->
    (ls, distributed[4N, 3+vd+1], indices[4N] i32, weights[4N])."""

    def __init__(self, experiment):
        super().__init__()
        self.experiment = experiment

    def forward(self, ls, positions, values, reset_hashmap=True):
        rows, idx, w = F_.DistributeLattice.apply(ls, positions, values, reset_hashmap,
                                                  self.experiment not in _NO_LOCAL_MEAN)
        return ls, rows, idx, w


class GroupNormLatticeModule(torch.nn.Module):
    """GroupNorm over [1,C,V]; parameters live in `gn` (a torch GroupNorm used as the container so
    the state-dict keys are gn.weight / gn.bias)."""

    def __init__(self, nr_params, affine=True):
        super().__init__()
        self.groups = ops.gn_groups(nr_params)
        self.gn = torch.nn.GroupNorm(self.groups, nr_params, affine=affine).to(_dev())

    def forward(self, lv, ls, relu=False):
        lv = ops.group_norm(lv, self.gn.weight, self.gn.bias, self.groups, self.gn.eps, relu)
        ls.set_values(lv)
        return lv, ls

    def affine(self, lv):
        """(sums, gamma, beta, eps) of this normalisation on `lv`, folded into the next fused kernel's gather"""
        return (ops.sums_of(lv, self.groups), self.gn.weight.detach(), self.gn.bias.detach(), self.gn.eps)

    def fusable(self, lv, nr_out):
        return (ops.no_grad_path(lv, self.gn.weight) and ops.conv_tc_supported(lv.shape[1], nr_out) and lv.shape[0] > 0
                and self.gn.weight is not None)


class Gn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.norm = None

    def forward(self, lv, ls):
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv.shape[1])
        return self.norm(lv, ls)


class Conv1x1(torch.nn.Module):
    """single tensor in / out (lattice_modules.py:95,102)"""

    def __init__(self, out_channels, bias):
        super().__init__()
        self.out_channels, self.use_bias, self.linear = out_channels, bias, None

    def forward(self, lv):
        if self.linear is None:
            self.linear = _linear(lv.shape[1], self.out_channels, self.use_bias)
        return ops.linear(lv, self.linear.weight, self.linear.bias)


class GnRelu1x1(torch.nn.Module):
    def __init__(self, out_channels, bias):
        super().__init__()
        self.out_channels, self.use_bias = out_channels, bias
        self.norm, self.linear = None, None

    def forward(self, lv, ls, res=None):
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv.shape[1])
            self.linear = _linear(lv.shape[1], self.out_channels, self.use_bias)
        if self.norm.fusable(lv, self.out_channels):  # GN + ReLU folded into the tensor-core kernel's A operand
            lv = ops.conv_tc(lv, None, ops.k_major(self.linear.weight, transposed=True), gn=self.norm.affine(lv), relu=True,
                             bias=None if self.linear.bias is None else self.linear.bias.detach(), res=res,
                             out_sums=ops.new_sums(self.out_channels, lv.device))
        elif F_.train_fusable(lv, lv.shape[1], self.out_channels, self.norm):   # training: the same fused kernel under autograd
            lv = F_.fused_conv_train(lv, self.linear.weight, self.linear.bias, res, self.norm, None, None, linear=True)
        else:
            lv, ls = self.norm(lv, ls, relu=True)
            lv = ops.linear(lv, self.linear.weight, self.linear.bias)
            if res is not None:
                lv = lv + res
        ls.set_values(lv)
        return lv, ls


class ConvLatticeModule(torch.nn.Module):
    """(lv[V,C], ls) -> (lv[V,F], ls); weight [9C, F], slot-major."""

    def __init__(self, nr_filters, neighbourhood_size=1, dilation=1, bias=True):
        super().__init__()
        self.nr_filters, self.neighbourhood_size, self.dilation, self.use_bias = nr_filters, neighbourhood_size, dilation, bias
        self.weight, self.bias = None, None

    def create(self, lv, ls):
        if self.weight is None:
            rows = ls.get_filter_extent(self.neighbourhood_size) * lv.shape[1]
            self.weight = _conv_weight(rows, self.nr_filters)
            if self.use_bias:
                b = 1.0 / math.sqrt(rows)
                self.bias = torch.nn.Parameter(torch.empty(self.nr_filters, device=_dev()).uniform_(-b, b))

    def forward(self, lv, ls, pre=None, res=None, norm=None):
        """pre = (sums, gamma, beta, eps): a GroupNorm+ReLU to fold into the gather (fused inference path only);
        norm: the GroupNormLatticeModule to fold in on the TRAINING path (funcs._FusedConv); res: residual added in the epilogue."""
        ls.set_values(lv)
        self.create(lv, ls)
        if ops.no_grad_path(lv, self.weight) and ops.conv_tc_supported(lv.shape[1], self.nr_filters) and lv.shape[0] > 0:
            out = ops.conv_tc(lv, ls.neighbours(dilation=self.dilation), ops.k_major(self.weight),
                              gn=pre, relu=pre is not None, bias=None if self.bias is None else self.bias.detach(), res=res,
                              out_sums=ops.new_sums(self.nr_filters, lv.device))
        elif norm is not None or (pre is None and F_.train_fusable(lv, lv.shape[1], self.nr_filters)):
            nbr = ls.neighbours(dilation=self.dilation)   # training: the same fused kernel under autograd (funcs._FusedConv)
            out = F_.fused_conv_train(lv, self.weight, self.bias, res, norm, nbr, lambda: nbr)
        else:
            if pre is not None:
                raise RuntimeError("a folded GroupNorm needs the fused kernel")
            out = F_.ConvIm2RowLattice.apply(lv, ls, self.weight, self.dilation)
            if self.bias is not None:
                out = out + self.bias
            if res is not None:
                out = out + res
        ls.set_values(out)
        return out, ls


class CoarsenLatticeModule(torch.nn.Module):
    def __init__(self, nr_filters):
        super().__init__()
        self.nr_filters, self.weight = nr_filters, None

    def forward(self, lv, ls, pre=None, norm=None):
        ls.set_values(lv)
        if self.weight is None:
            self.weight = _conv_weight(ls.get_filter_extent(1) * lv.shape[1], self.nr_filters)
        if ops.no_grad_path(lv, self.weight) and ops.conv_tc_supported(lv.shape[1], self.nr_filters) and lv.shape[0] > 0:
            coarse = ls.create_coarse_verts()
            out = ops.conv_tc(lv, coarse.neighbours(ls, mode=1), ops.k_major(self.weight),
                              gn=pre, relu=pre is not None, out_sums=ops.new_sums(self.nr_filters, lv.device))
        elif norm is not None or (pre is None and F_.train_fusable(lv, lv.shape[1], self.nr_filters)):
            coarse = ls.create_coarse_verts()
            out = F_.fused_conv_train(lv, self.weight, None, None, norm, coarse.neighbours(ls, mode=1),
                                      lambda: ls.neighbours(coarse, mode=2))
        else:
            if pre is not None:
                raise RuntimeError("a folded GroupNorm needs the fused kernel")
            out, coarse = F_.CoarsenLattice.apply(lv, ls, self.weight)
        coarse.set_values(out)
        return out, coarse


class FinefyLatticeModule(torch.nn.Module):
    def __init__(self, nr_filters):
        super().__init__()
        self.nr_filters, self.weight = nr_filters, None

    def forward(self, lv_coarse, ls_coarse, ls_fine, pre=None, norm=None):
        ls_coarse.set_values(lv_coarse)
        if self.weight is None:
            self.weight = _conv_weight(ls_coarse.get_filter_extent(1) * lv_coarse.shape[1], self.nr_filters)
        if (ops.no_grad_path(lv_coarse, self.weight) and ops.conv_tc_supported(lv_coarse.shape[1], self.nr_filters)
                and lv_coarse.shape[0] > 0):
            out = ops.conv_tc(lv_coarse, ls_fine.neighbours(ls_coarse, mode=2), ops.k_major(self.weight),
                              gn=pre, relu=pre is not None)
        elif norm is not None or (pre is None and F_.train_fusable(lv_coarse, lv_coarse.shape[1], self.nr_filters)):
            out = F_.fused_conv_train(lv_coarse, self.weight, None, None, norm, ls_fine.neighbours(ls_coarse, mode=2),
                                      lambda: ls_coarse.neighbours(ls_fine, mode=1))
        else:
            if pre is not None:
                raise RuntimeError("a folded GroupNorm needs the fused kernel")
            out = F_.FinefyLattice.apply(lv_coarse, ls_coarse, ls_fine, self.weight)
        ls_fine.set_values(out)
        return out, ls_fine


class GnReluConv(torch.nn.Module):
    def __init__(self, nr_filters, dilation, bias, with_dropout):
        super().__init__()
        self.norm = None
        self.conv = ConvLatticeModule(nr_filters, 1, dilation, bias)
        self.drop = torch.nn.Dropout(0.2) if with_dropout else None

    def forward(self, lv, ls, res=None):
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv.shape[1])
        if self.norm.fusable(lv, self.conv.nr_filters) and (self.drop is None or not self.training):
            return self.conv(lv, ls, pre=self.norm.affine(lv), res=res)
        if F_.train_fusable(lv, lv.shape[1], self.conv.nr_filters, self.norm) and (self.drop is None or not self.training):
            return self.conv(lv, ls, res=res, norm=self.norm)
        lv, ls = self.norm(lv, ls, relu=True)
        if self.drop is not None:
            lv = self.drop(lv)
        return self.conv(lv, ls, res=res)


class GnReluCoarsen(torch.nn.Module):
    """models.py:182,353: (lv_fine, ls_fine) -> (lv_coarse[Vc,F], ls_coarse)"""

    def __init__(self, nr_filters):
        super().__init__()
        self.norm = None
        self.coarse = CoarsenLatticeModule(nr_filters)

    def forward(self, lv, ls):
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv.shape[1])
        if self.norm.fusable(lv, self.coarse.nr_filters):
            return self.coarse(lv, ls, pre=self.norm.affine(lv))
        if F_.train_fusable(lv, lv.shape[1], self.coarse.nr_filters, self.norm):
            return self.coarse(lv, ls, norm=self.norm)
        lv, ls = self.norm(lv, ls, relu=True)
        return self.coarse(lv, ls)


class GnReluFinefy(torch.nn.Module):
    """models.py:214,398: (lv_coarse, ls_coarse, ls_fine) -> (lv_fine[Vf,F], ls_fine)"""

    def __init__(self, nr_filters):
        super().__init__()
        self.norm = None
        self.fine = FinefyLatticeModule(nr_filters)

    def forward(self, lv_coarse, ls_coarse, ls_fine):
        if self.norm is None:
            self.norm = GroupNormLatticeModule(lv_coarse.shape[1])
        if self.norm.fusable(lv_coarse, self.fine.nr_filters):
            return self.fine(lv_coarse, ls_coarse, ls_fine, pre=self.norm.affine(lv_coarse))
        if F_.train_fusable(lv_coarse, lv_coarse.shape[1], self.fine.nr_filters, self.norm):
            return self.fine(lv_coarse, ls_coarse, ls_fine, norm=self.norm)
        lv_coarse, ls_coarse = self.norm(lv_coarse, ls_coarse, relu=True)
        return self.fine(lv_coarse, ls_coarse, ls_fine)


class ResnetBlock(torch.nn.Module):
    """models.py:175,227: 2 x (GN -> ReLU -> conv) + identity"""

    def __init__(self, nr_filters, dilations, biases, with_dropout):
        super().__init__()
        self.conv1 = GnReluConv(nr_filters, dilations[0], biases[0], False)
        self.conv2 = GnReluConv(nr_filters, dilations[1], biases[1], with_dropout)

    def forward(self, lv, ls):
        skip = lv
        lv, ls = self.conv1(lv, ls)
        lv, ls = self.conv2(lv, ls, res=skip)  # identity added in the conv epilogue
        ls.set_values(lv)
        return lv, ls


class BottleneckBlock(torch.nn.Module):
    """models.py:178,193,230: 1x1 (C/4) -> conv (C/4) -> 1x1 (C) + identity"""

    def __init__(self, out_channels, biases):
        super().__init__()
        self.contract = GnRelu1x1(int(out_channels / 4), biases[0])
        self.conv = GnReluConv(int(out_channels / 4), 1, biases[1], False)
        self.expand = GnRelu1x1(out_channels, biases[2])

    def forward(self, lv, ls):
        skip = lv
        lv, ls = self.contract(lv, ls)
        lv, ls = self.conv(lv, ls)
        lv, ls = self.expand(lv, ls, res=skip)
        ls.set_values(lv)
        return lv, ls


class SliceLatticeModule(torch.nn.Module):
    def forward(self, lv, ls, positions, indices=None, weights=None):
        ls.set_values(lv)
        return F_.SliceLattice.apply(lv, ls, positions, indices, weights)


class SplatLatticeModule(torch.nn.Module):
    def forward(self, ls, positions, values):
        lv, idx, w = F_.SplatLattice.apply(ls, positions, values)
        ls.set_values(lv)
        return lv, ls, idx, w


class SliceFastCUDALatticeModule(torch.nn.Module):
    """models.py:232,465 (recipe SURVEY E12): 1x1 chain C -> C -> C/2 -> 8, gather [N, 4*9], subtract
    gamma*max-over-simplex + beta, Linear 36->36, GN, ReLU, Linear 36->4 = delta weights,
    slice_classify(lv, delta, W[classes, C], b)."""

    def __init__(self, nr_classes, dropout_prob, experiment):
        super().__init__()
        self.nr_classes, self.experiment = nr_classes, experiment
        self.bottleneck_size = 8
        self.stepdown = torch.nn.ModuleList([])
        self.bottleneck = None
        self.linear_pre_deltaW = None
        self.dropout = torch.nn.Dropout(dropout_prob) if dropout_prob > 0.0 else None

    def _create(self, C):
        dev = _dev()
        for i in range(2):
            self.stepdown.append(GnRelu1x1(int(C / (2 ** i)), False))
        self.bottleneck = GnRelu1x1(self.bottleneck_size, False)
        g = 4 * (self.bottleneck_size + 1)
        self.linear_pre_deltaW = torch.nn.Linear(g, g, bias=False).to(dev)
        self.gn_middle = GroupNormLatticeModule(g)
        self.linear_deltaW = torch.nn.Linear(g, 4, bias=True).to(dev)
        self.linear_clasify = torch.nn.Linear(C, self.nr_classes, bias=True).to(dev)
        self.gamma = torch.nn.Parameter(torch.ones(self.bottleneck_size + 1, device=dev))
        self.beta = torch.nn.Parameter(torch.zeros(self.bottleneck_size + 1, device=dev))
        with torch.no_grad():
            torch.nn.init.kaiming_uniform_(self.linear_pre_deltaW.weight, mode="fan_in", nonlinearity="relu")
            self.linear_deltaW.weight.mul_(0.1)
            self.linear_deltaW.bias.zero_()

    def forward(self, lv, ls, positions, indices, weights):
        if self.bottleneck is None:
            self._create(lv.shape[1])
        if self.dropout is not None:
            lv = self.dropout(lv)
        ls.set_values(lv)
        b, lsb = lv, ls
        for i in range(2):
            b, lsb = self.stepdown[i](b, lsb)
        b, lsb = self.bottleneck(b, lsb)
        W, cb = self.linear_clasify.weight, self.linear_clasify.bias
        N = positions.shape[0]
        if (ops.no_grad_path(lv, b, W, self.linear_pre_deltaW.weight) and lv.shape[1] % 32 == 0 and lv.shape[0] > 0 and N > 0
                and b.shape[1] == 8 and self.nr_classes <= 32 and self.gn_middle.gn.weight is not None):
            # inference: classify the VERTICES once (slicing is linear: a [V,C] x [C,classes] tensor-core GEMM, classes padded to
            # a multiple of 8), then ONE pair of kernels does everything per point -- gather, max over the simplex, both
            # small linears with the GroupNorm between them, the slice of the class scores with the deformed weights, bias and
            # the model's log-softmax (csrc/ltn_slice_head.cu)
            kp = (self.nr_classes + 7) // 8 * 8
            ls.set_values(lv)
            scores = ops.conv_tc(lv, None, ops.k_major_padded(W, kp))
            logits = torch.empty(N, self.nr_classes, dtype=torch.float32, device=lv.device)
            logsm = torch.empty_like(logits)
            sums = torch.empty(18, 2, dtype=torch.float64, device=lv.device)
            p = _lib.ptr
            rc = _lib.load().ltn_slice_head(p(b.contiguous()), b.shape[0], _lib.rows_dev(b.shape[0]), p(scores), scores.stride(0),
                                            p(indices), p(weights), N, _lib.rows_dev(N), p(self.gamma.detach()), p(self.beta.detach()),
                                            p(self.linear_pre_deltaW.weight.detach()), p(self.gn_middle.gn.weight.detach()),
                                            p(self.gn_middle.gn.bias.detach()), float(self.gn_middle.gn.eps),
                                            p(self.linear_deltaW.weight.detach()), p(self.linear_deltaW.bias.detach()),
                                            None if cb is None else p(cb.detach()), self.nr_classes,
                                            1 if self.experiment == "slice_no_deform" else 0, p(sums), p(logits), p(logsm), _lib.stream())
            _lib.check(rc, "ltn_slice_head")
            logits._ltn_logsoftmax = logsm   # LatticeNetSeq.forward returns it instead of running LogSoftmax again
            return logits
        g = F_.GatherLattice.apply(b, lsb, positions, indices, weights)
        g3 = g.view(N, 4, self.bottleneck_size + 1)
        mx = g3.max(1, keepdim=True)[0]
        g = (g3 - (self.gamma * mx + self.beta)).reshape(N, -1)
        g = ops.linear(g, self.linear_pre_deltaW.weight)
        g, _ = self.gn_middle(g, lsb, relu=True)
        dw = ops.linear(g, self.linear_deltaW.weight, self.linear_deltaW.bias)
        if self.experiment == "slice_no_deform":
            dw = dw * 0
        ls.set_values(lv)
        W, b = self.linear_clasify.weight, self.linear_clasify.bias
        if ops.no_grad_path(lv, dw, W) and lv.shape[1] % 32 == 0 and lv.shape[0] > 0:
            # slicing is linear, so classify the VERTICES once (a [V,C] x [C,classes] tensor-core GEMM, classes padded to
            # a multiple of 8) and slice the class scores: logit[p] = b + sum_r (w+dw)[p,r] * (lv[idx] @ W^T) --
            # 4 x 128-byte gathers per point instead of 4 x C floats and a per-point mat-vec
            kp = (self.nr_classes + 7) // 8 * 8
            scores = ops.conv_tc(lv, None, ops.k_major_padded(W, kp))
            ww = (weights + dw.reshape(-1)).contiguous()
            sliced = F_.SliceLattice.apply(scores, ls, positions, indices, ww)
            return sliced[:, : self.nr_classes] + b.detach()
        return F_.SliceClassifyLattice.apply(lv, ls, positions, dw, W, b, self.nr_classes, indices, weights)
