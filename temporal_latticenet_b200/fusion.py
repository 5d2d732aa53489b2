"""Per-vertex temporal fusion modules on the sm_100a kernels.

Behavioural mirror of /root/reference/seq_lattice/lattice_modules.py:17-339 (LSTMModule, GRUModule,
CrossframeGlobalAttentionModule, TemporalMaxPoolModule, TemporalLinearModule,
CrossframeLocalInterpolationModule + CustomKernelConvLatticeIm2RowModule): same class names,
constructor arguments, parameter names (state-dicts interchange) and the same index-prefix
alignment of the hidden state (vertex ids are append-only within a window, so h^{t-1} covers the
first V_{t-1} rows of frame t and the rest is padding).

What differs is the execution: the padding is never materialised, the GRU/LSTM gates are one
pointwise kernel after the two gate GEMMs, and AFlow's gather + distance + weights + weighted sum is
a single kernel (csrc/ltn_fusion.cu).  With autograd enabled the same math runs as differentiable
torch ops over the im2row Function so BPTT through the 4 frames works (lattice_modules.py:56-63).
"""
import math

import torch

from . import _lib
from . import ops
from .funcs import Im2RowLattice
from .modules import Conv1x1, Gn

AFLOW_PAD = -999999.0  # lattice_modules.py:215


def _pad_rows(h, nr_rows, value=0.0):
    if h.shape[0] > nr_rows:
        raise RuntimeError("hidden state has more vertices (%d) than the current lattice values (%d): "
                           "vertex ids must be append-only within a window" % (h.shape[0], nr_rows))
    if h.shape[0] == nr_rows:
        return h
    return torch.nn.functional.pad(h, (0, 0, 0, nr_rows - h.shape[0]), value=value)


def _fused_ok(*tensors):
    return not (torch.is_grad_enabled() and any(t.requires_grad for t in tensors if t is not None))


class _HiddenState:
    """Hidden state h^{t-1} of a fusion module.  Eager mode: a reference to the previous frame's tensor.
    Static-capacity mode (engine.py): a persistent [capacity, C] buffer plus the device-side row count of
    the frame that wrote it, so every captured frame graph reads and writes the same addresses."""

    def _store(self, lv):
        if _lib.static_mode():
            if getattr(self, "_h_buf", None) is None or self._h_buf.shape != lv.shape:
                self._h_buf = torch.empty_like(lv)
                self._h_rows = torch.zeros(1, dtype=torch.int32, device=lv.device)
            if lv.data_ptr() != self._h_buf.data_ptr():   # a module that wrote its output into _out_buffer() skips the copy
                self._h_buf.copy_(lv)
            self._h_rows.copy_(_lib.rows_tensor(lv.shape[0]))
            self.h_lv = self._h_buf
        else:
            self.h_lv = lv

    def _out_buffer(self, like):
        """where a module whose inputs are all OTHER tensors (the gates, the projected hidden state) may write its output
        directly: static-capacity mode -> the persistent hidden-state buffer itself (the frame's later layers read it there,
        the next frame finds it as h^{t-1}: one [V, C] copy per fusion point and frame less); eager mode -> a fresh tensor"""
        buf = getattr(self, "_h_buf", None)
        if _lib.static_mode() and buf is not None and buf.shape == like.shape:
            return buf
        return torch.empty_like(like)

    def _rows_dev(self):
        if _lib.static_mode() and getattr(self, "_h_rows", None) is not None:
            return _lib.ptr(self._h_rows)
        return None

    def _has_history(self, nr_rows):
        """static-capacity mode: bool [capacity, 1], true for the rows the hidden state really covers (the rows of the frame
        that wrote it); the buffer beyond them is stale and stands for the reference's padding"""
        rows = torch.arange(nr_rows, dtype=torch.int32, device=self._h_rows.device)
        return (rows < self._h_rows).unsqueeze(1)


class GRUModule(torch.nn.Module, _HiddenState):
    """lattice_modules.py:42-66"""

    def __init__(self, nr_output_channels):
        super().__init__()
        self.GRU = torch.nn.GRUCell(nr_output_channels, nr_output_channels, bias=True)  # parameter container
        self.hidden_linear = torch.nn.Linear(nr_output_channels, nr_output_channels)
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None

    def forward(self, lv, ls):
        if self.h_lv is None:
            self._store(lv)
            return lv, ls
        g = self.GRU
        hrows = self._rows_dev()   # static mode: h covers the PREVIOUS frame's rows; the buffer beyond them is stale
        h = ops.linear(self.h_lv, self.hidden_linear.weight, self.hidden_linear.bias, rows_dev=hrows)
        V, C, Vh = lv.shape[0], lv.shape[1], h.shape[0]
        if _fused_ok(lv, h, g.weight_ih):
            if Vh > V:
                _pad_rows(h, V)
            gi = ops.linear(lv, g.weight_ih, g.bias_ih)
            gh = ops.linear(h, g.weight_hh, g.bias_hh, rows_dev=hrows)
            new_lv = self._out_buffer(lv)   # h (the projection of h^{t-1}), gi and gh are separate tensors: writing in place is safe
            p = _lib.ptr
            if C % 4 == 0 and C <= 256:
                # the statistics the next layer's GroupNorm needs come out of the same pass (no k_gn_stats read of h')
                sums = ops.new_sums(C, lv.device)
                _lib.check(_lib.load().ltn_gru_pointwise_stats(p(gi), p(gh), p(h.contiguous()), p(g.bias_hh), V, Vh, _lib.rows_dev(V),
                                                               self._rows_dev(), C, p(new_lv), p(sums), sums.shape[0], _lib.stream()),
                           "ltn_gru_pointwise_stats")
                new_lv._ltn_gn_sums = (ops._FRAME["id"], sums)
            else:
                _lib.check(_lib.load().ltn_gru_pointwise(p(gi), p(gh), p(h.contiguous()), p(g.bias_hh), V, Vh, _lib.rows_dev(V),
                                                         self._rows_dev(), C, p(new_lv), _lib.stream()), "ltn_gru_pointwise")
        else:
            new_lv = g(lv, _pad_rows(h, V))
        self._store(new_lv)
        ls.set_values(new_lv)
        return new_lv, ls


class LSTMModule(torch.nn.Module, _HiddenState):
    """lattice_modules.py:17-40 (the cell state is discarded: c_prev = 0 every frame)"""

    def __init__(self, nr_output_channels):
        super().__init__()
        self.lstm = torch.nn.LSTMCell(nr_output_channels, nr_output_channels, bias=True)
        self.hidden_linear = torch.nn.Linear(nr_output_channels, nr_output_channels)
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None

    def forward(self, lv, ls):
        if self.h_lv is None:
            self._store(lv)
            return lv, ls
        c = self.lstm
        hrows = self._rows_dev()
        h = ops.linear(self.h_lv, self.hidden_linear.weight, self.hidden_linear.bias, rows_dev=hrows)
        V, C, Vh = lv.shape[0], lv.shape[1], h.shape[0]
        if _fused_ok(lv, h, c.weight_ih):
            if Vh > V:
                _pad_rows(h, V)
            gi = ops.linear(lv, c.weight_ih, c.bias_ih)
            gh = ops.linear(h, c.weight_hh, c.bias_hh, rows_dev=hrows)
            new_lv = torch.empty_like(lv)
            p = _lib.ptr
            _lib.check(_lib.load().ltn_lstm_pointwise(p(gi), p(gh), p(c.bias_hh), V, Vh, _lib.rows_dev(V), self._rows_dev(), C,
                                                      p(new_lv), _lib.stream()), "ltn_lstm_pointwise")
        else:
            hp = _pad_rows(h, V)
            new_lv, _ = c(lv, (hp, torch.zeros_like(hp)))
        self._store(new_lv)
        ls.set_values(new_lv)
        return new_lv, ls


class CrossframeGlobalAttentionModule(torch.nn.Module, _HiddenState):
    """lattice_modules.py:70-116 (quirk Q6: the same 1x1 conv twice, "pooling" = 1/(rows+cols))"""

    def __init__(self, nr_output_channels):
        super().__init__()
        self.groupnorm = Gn()
        self.conv = Conv1x1(out_channels=nr_output_channels, bias=False)
        self.hidden_linear = torch.nn.Linear(nr_output_channels, nr_output_channels)
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None

    def forward(self, lv, ls):
        if self.h_lv is None:
            self._store(lv)
            return lv, ls
        static = _lib.static_mode()
        h = ops.linear(self.h_lv, self.hidden_linear.weight, self.hidden_linear.bias, rows_dev=self._rows_dev())
        Vh, V = h.shape[0], lv.shape[0]
        if static:   # same row capacity; the live counts are on the device
            known = self._has_history(V)
            hp = torch.where(known, h, torch.zeros_like(h))
            nr_rows = _lib.rows_tensor(V).to(torch.float32)
        else:
            hp = _pad_rows(h, V)
        a = torch.relu(self.conv(hp))
        a, _ = self.groupnorm(a, ls)
        a = self.conv(a)
        if static:
            a = torch.sigmoid(a * (1.0 / (nr_rows + float(a.shape[1]))))
            a = torch.where(known, a, torch.ones_like(a))
        else:
            a = torch.sigmoid(a * (1.0 / (a.shape[0] + a.shape[1])))
            if Vh < V:
                a = torch.cat([a[:Vh], torch.ones(V - Vh, a.shape[1], dtype=a.dtype, device=a.device)], 0)
        lv = a * lv
        self._store(lv)
        ls.set_values(lv)
        return lv, ls


class TemporalMaxPoolModule(torch.nn.Module, _HiddenState):
    """lattice_modules.py:119-145"""

    def __init__(self):
        super().__init__()
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None

    def forward(self, lv, ls):
        if self.h_lv is None:
            self._store(lv)
        elif _lib.static_mode():
            hp = torch.where(self._has_history(lv.shape[0]), self.h_lv, torch.full_like(self.h_lv, -9999.0))
            lv = torch.maximum(hp, lv)
            self._store(lv)
        else:
            n = max(self.h_lv.shape[0], lv.shape[0])
            hp = _pad_rows(self.h_lv, n, -9999.0)
            lv = torch.maximum(hp, _pad_rows(lv, n, -9999.0))
            self.h_lv = 0.0 * hp + lv
        ls.set_values(lv)
        return lv, ls


class TemporalLinearModule(torch.nn.Module, _HiddenState):
    """lattice_modules.py:149-185"""

    def __init__(self, nr_output_channels):
        super().__init__()
        self.nr_output_channels = nr_output_channels
        self.linear = torch.nn.Linear(nr_output_channels * 2, nr_output_channels)
        self.hidden_linear = torch.nn.Linear(nr_output_channels, nr_output_channels)
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None

    def forward(self, lv, ls):
        if self.h_lv is None:
            if lv.shape[1] != self.nr_output_channels:
                raise RuntimeError("lv has %d channels, the module was built for %d" % (lv.shape[1], self.nr_output_channels))
            self._store(lv)
        else:
            h = ops.linear(self.h_lv, self.hidden_linear.weight, self.hidden_linear.bias, rows_dev=self._rows_dev())
            if _lib.static_mode():
                hp = torch.where(self._has_history(lv.shape[0]), h, torch.zeros_like(h))
                lv = torch.relu(ops.linear(torch.cat([hp, lv], 1), self.linear.weight, self.linear.bias))
            else:
                hp = _pad_rows(h, lv.shape[0])
                lv = 0.0 * hp + torch.relu(ops.linear(torch.cat([hp, lv], 1), self.linear.weight, self.linear.bias))
            self._store(lv)
        ls.set_values(lv)
        return lv, ls


class CustomKernelConvLatticeIm2RowModule(torch.nn.Module):
    """AFlow core, lattice_modules.py:238-339.  `weight` [9C, C] is created (and lives in the
    state-dict) but takes no part in the computation -- quirk Q4."""

    def __init__(self, nr_filters, neighbourhood_size=1, dilation=1, bias=True, use_center=True, train_alpha_beta=True):
        super().__init__()
        self.nr_filters, self.neighbourhood_size, self.dilation = nr_filters, neighbourhood_size, dilation
        self.use_bias, self.use_center = bias, use_center
        self.weight, self.bias = None, None
        if train_alpha_beta:
            self.alpha = torch.nn.Parameter(torch.tensor(0.1))
            self.beta = torch.nn.Parameter(torch.tensor(0.1))
        else:
            self.register_buffer("alpha", torch.tensor(0.1), persistent=False)
            self.register_buffer("beta", torch.tensor(0.1), persistent=False)

    def _create(self, ls, val_dim, device):
        rows = ls.get_filter_extent(self.neighbourhood_size) * val_dim
        bound = math.sqrt(3.0) * math.sqrt(2.0) / math.sqrt(rows)
        self.weight = torch.nn.Parameter(torch.empty(rows, self.nr_filters, device=device).uniform_(-bound, bound))
        if self.use_bias:
            b = 1.0 / math.sqrt(rows)
            self.bias = torch.nn.Parameter(torch.empty(self.nr_filters, device=device).uniform_(-b, b))

    def forward(self, lattice_values, hidden_state, lattice_structure, nr_hidden_rows=None, hidden_rows_dev=None):
        """hidden_state: h^{t-1}, either already padded to V rows (reference call shape) or the
        unpadded [Vh,C] tensor with nr_hidden_rows=None meaning "all rows are real"."""
        ls, lv = lattice_structure, lattice_values
        ls.set_values(lv)
        if self.weight is None:
            self._create(ls, lv.shape[1], lv.device)
        V, C = lv.shape
        nbr = ls.neighbours(dilation=1)
        Vh = hidden_state.shape[0] if nr_hidden_rows is None else nr_hidden_rows
        if _fused_ok(lv, hidden_state, self.alpha):
            out = torch.empty_like(lv)
            weights = torch.empty(V, 9, dtype=torch.float32, device=lv.device)
            p = _lib.ptr
            _lib.check(_lib.load().ltn_aflow(p(lv.contiguous()), p(hidden_state.contiguous()), V, Vh, _lib.rows_dev(V),
                                             hidden_rows_dev, C, p(nbr),
                                             p(self.alpha.detach().reshape(1)), p(self.beta.detach().reshape(1)),
                                             p(self.bias) if self.bias is not None else None, AFLOW_PAD,
                                             1 if self.use_center else 0, p(out), p(weights), _lib.stream()), "ltn_aflow")
        else:
            hp = _pad_rows(hidden_state[:Vh], V, AFLOW_PAD)
            nb = Im2RowLattice.apply(hp, ls, 9, 1, self.nr_filters).reshape(V, 9, C)
            present = (nbr != -1).to(lv.dtype)
            dist = torch.cdist(nb, lv.unsqueeze(1), p=2.0).squeeze(2) * present
            if not self.use_center:
                dist = torch.cat([dist[:, :-1], dist[:, -1:] * 0.0], 1)
            dist = dist / dist.sum(1, keepdim=True).detach()
            alpha_t = torch.ones_like(dist) * self.alpha
            weights = (alpha_t - torch.min(dist, alpha_t)) * self.beta * present
            if not self.use_center:
                weights = torch.cat([weights[:, :-1], weights[:, -1:] * 0.0], 1)
            out = (nb * weights.unsqueeze(2)).sum(1)
            if self.bias is not None:
                out = out + self.bias
        ls.set_values(lv)
        return out, weights, nbr


class CrossframeLocalInterpolationModule(torch.nn.Module, _HiddenState):
    """AFlow wrapper, lattice_modules.py:188-235"""

    def __init__(self, nr_output_channels, train_alpha_beta=True, use_center=True):
        super().__init__()
        self.nr_output_channels = nr_output_channels
        self.AFLOW = CustomKernelConvLatticeIm2RowModule(nr_filters=nr_output_channels, train_alpha_beta=train_alpha_beta,
                                                         use_center=use_center)
        self.linear = torch.nn.Linear(nr_output_channels * 2, nr_output_channels)
        self.h_lv = None
        self.h_lv_vis, self.weights_vis, self.lattice_neighbors_previous = None, None, None

    def reset_sequence(self):
        self.h_lv = None
        self.h_lv_vis, self.weights_vis, self.lattice_neighbors_previous = None, None, None

    def return_for_vis(self):
        return self.h_lv_vis, self.weights_vis, self.lattice_neighbors_previous

    def forward(self, lv, ls):
        if self.h_lv is None:
            self._store(lv)
        else:
            feat, weights, nbr = self.AFLOW(lv, self.h_lv, ls, hidden_rows_dev=self._rows_dev())
            self.h_lv_vis, self.weights_vis, self.lattice_neighbors_previous = self.h_lv.detach(), weights.detach(), nbr
            lv = torch.relu(ops.linear(torch.cat([feat, lv], 1), self.linear.weight, self.linear.bias))
            self._store(lv)
        ls.set_values(lv)
        return lv, ls


FUSION_TYPES = ("linear", "maxpool", "cga", "aflow", "lstm", "gru")


def make_fusion(kind, nr_channels):
    """models.py:76-153 / lattice_modules.py:364-386 dispatch"""
    if kind == "linear": return TemporalLinearModule(nr_channels)
    if kind == "maxpool": return TemporalMaxPoolModule()
    if kind == "cga": return CrossframeGlobalAttentionModule(nr_channels)
    if kind == "aflow": return CrossframeLocalInterpolationModule(nr_channels)
    if kind == "lstm": return LSTMModule(nr_channels)
    if kind == "gru": return GRUModule(nr_channels)
    return None
