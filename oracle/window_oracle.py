"""CPU oracle of the whole window: the reference's model recipe (seq_lattice/models.py::LNN_SEQ,
seq_lattice/lattice_modules.py) restated as plain, unfused torch-CPU code over the oracle shims
(oracle/shims: scalar C lattice core + torch indexing).

TEST INFRASTRUCTURE ONLY (part of oracle/): the checker for __graft_entry__.smoke() and the timed
`cpu_baseline` / `--impl reference` legs of bench.py on the GPU box, where /root/reference does not
exist.  PARITY: pinned against tests/golden/*.npz, which were produced by the reference's OWN
unmodified models.py / lattice_modules.py running over the same shims (tests/test_oracle_cpu.py);
at the `latticenet` boundary below the shims parity stays UNPINNED (see oracle/lattice_oracle.c).

Every class cites the reference lines it follows; quirks Q1-Q8 of SURVEY.md appendix A are kept.
"""
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

import hjson  # noqa: E402
import latticenet  # noqa: E402
if not os.path.abspath(latticenet.__file__).startswith(_SHIMS):
    raise RuntimeError("oracle/window_oracle.py needs the ORACLE shims first on sys.path, found %s" % latticenet.__file__)
import torch_scatter  # noqa: E402
from latticenet import Lattice, ModelParams  # noqa: E402
from latticenet_py.lattice import lattice_funcs as LF  # noqa: E402
from latticenet_py.lattice import lattice_modules as LM  # noqa: E402

_KINDS = ("linear", "maxpool", "cga", "aflow", "lstm", "gru")


def _pad(h, rows, value):
    return torch.nn.functional.pad(h, (0, 0, 0, rows - h.shape[0]), value=value)


class _Recurrent(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.h_lv = None

    def reset_sequence(self):
        self.h_lv = None


class LSTMModule(_Recurrent):
    """lattice_modules.py:17-40"""

    def __init__(self, C):
        super().__init__()
        self.lstm = torch.nn.LSTMCell(C, C, bias=True)
        self.hidden_linear = torch.nn.Linear(C, C)

    def forward(self, lv, ls):
        if self.h_lv is None:
            self.h_lv = lv.clone()
        else:
            h = _pad(self.hidden_linear(self.h_lv), lv.shape[0], 0.0)        # :32-34
            lv, _ = self.lstm(lv, (h, torch.zeros_like(h)))                   # :36 (cell state dropped)
            self.h_lv = lv.clone()
            ls.set_values(lv)
        return lv, ls


class GRUModule(_Recurrent):
    """lattice_modules.py:42-66"""

    def __init__(self, C):
        super().__init__()
        self.GRU = torch.nn.GRUCell(C, C, bias=True)
        self.hidden_linear = torch.nn.Linear(C, C)

    def forward(self, lv, ls):
        if self.h_lv is None:
            self.h_lv = lv.clone()
            return lv.clone(), ls
        h = _pad(self.hidden_linear(self.h_lv), lv.shape[0], 0.0)            # :58-60
        new_lv = self.GRU(lv, h)                                              # :62
        self.h_lv = new_lv.clone()
        ls.set_values(new_lv)
        return new_lv, ls


class CrossframeGlobalAttentionModule(_Recurrent):
    """lattice_modules.py:70-116 (Q6: same 1x1 conv twice; "pooling" = 1/(rows+cols))"""

    def __init__(self, C):
        super().__init__()
        self.groupnorm = LM.Gn()
        self.conv = LM.Conv1x1(out_channels=C, bias=False)
        self.hidden_linear = torch.nn.Linear(C, C)

    def forward(self, lv, ls):
        if self.h_lv is None:
            self.h_lv = lv.clone()
            return lv, ls
        prev_rows = self.h_lv.shape[0]
        h = _pad(self.hidden_linear(self.h_lv), lv.shape[0], 0.0)            # :90-92
        h = torch.relu(self.conv(h))                                          # :95-98
        h, _ = self.groupnorm(h, ls)                                          # :100
        h = self.conv(h)                                                      # :102
        h = torch.sigmoid(h * torch.tensor(1 / (h.shape[0] + h.shape[1])))   # :104-106
        if prev_rows < lv.shape[0]:                                           # :109-110 (one-padding)
            h = torch.cat([h[:prev_rows], torch.ones(lv.shape[0] - prev_rows, h.shape[1])], 0)
        lv = h * lv
        self.h_lv = lv.clone()
        ls.set_values(lv)
        return lv, ls


class TemporalMaxPoolModule(_Recurrent):
    """lattice_modules.py:119-145"""

    def forward(self, lv, ls):
        if self.h_lv is None:
            self.h_lv = lv.clone()
        else:
            padded = torch.nn.utils.rnn.pad_sequence([self.h_lv, lv], padding_value=-9999.0)  # :137
            h = padded[:, 0, :]
            lv, _ = torch.max(padded, dim=1)                                  # :140
            self.h_lv = 0.0 * h + lv.clone()                                  # :141 (alpha = 0)
        ls.set_values(lv)
        return lv, ls


class TemporalLinearModule(_Recurrent):
    """lattice_modules.py:149-185"""

    def __init__(self, C):
        super().__init__()
        self.nr_output_channels = C
        self.linear = torch.nn.Linear(2 * C, C)
        self.hidden_linear = torch.nn.Linear(C, C)

    def forward(self, lv, ls):
        if self.h_lv is None:
            if lv.shape[1] != self.nr_output_channels:
                raise RuntimeError("channel mismatch (lattice_modules.py:165-167)")
            self.h_lv = lv.clone()
        else:
            h = _pad(self.hidden_linear(self.h_lv), lv.shape[0], 0.0)        # :171-174
            cat = torch.relu(self.linear(torch.cat([h, lv], 1)))             # :175-178
            lv = 0.0 * h + cat                                                # :180
            self.h_lv = lv.clone()
        ls.set_values(lv)
        return lv, ls


class CustomKernelConvLatticeIm2RowModule(torch.nn.Module):
    """AFlow core, lattice_modules.py:238-339 (Q4: `weight` exists but is unused; Q5: 0/0 kept)"""

    def __init__(self, nr_filters, use_center=True):
        super().__init__()
        self.nr_filters, self.use_center = nr_filters, use_center
        self.weight, self.bias = None, None
        self.alpha = torch.nn.Parameter(torch.tensor(0.1))
        self.beta = torch.nn.Parameter(torch.tensor(0.1))

    def forward(self, lv, hidden, ls):
        ls.set_values(lv)
        fe = ls.get_filter_extent(1)
        if self.weight is None:
            self.weight = LM._conv_weight(fe * ls.val_dim(), self.nr_filters)                      # :291
            b = 1.0 / (fe * ls.val_dim()) ** 0.5
            self.bias = torch.nn.Parameter(torch.empty(self.nr_filters).uniform_(-b, b))           # :292-295
        C, V = self.nr_filters, lv.shape[0]
        ls.set_values(hidden)
        nb = LF.Im2RowLattice.apply(hidden, ls, fe, 1, C).reshape(V, -1, C)                        # :300-301
        idx = LF.Im2RowIndicesLattice.apply(hidden, ls, fe, 1, C)[:, ::C]                          # :304,318
        present = (idx != -1)
        d = torch.cdist(nb, lv.unsqueeze(1), p=2.0).squeeze(2) * present                           # :316-318
        if not self.use_center:
            d[:, -1] = d[:, -1] * 0.0
        d = d * 1 / d.sum(1, keepdim=True).detach()                                                 # :321
        a = torch.ones_like(d) * self.alpha
        w = (a - torch.min(d, a)) * self.beta * present                                            # :324-325
        if not self.use_center:
            w[:, -1] = w[:, -1] * 0.0
        out = (nb.permute(0, 2, 1) * w.unsqueeze(1)).sum(2) + self.bias                            # :331-334
        ls.set_values(lv)
        return out, w, idx


class CrossframeLocalInterpolationModule(_Recurrent):
    """AFlow wrapper, lattice_modules.py:188-235"""

    def __init__(self, C):
        super().__init__()
        self.AFLOW = CustomKernelConvLatticeIm2RowModule(C)
        self.linear = torch.nn.Linear(2 * C, C)

    def forward(self, lv, ls):
        if self.h_lv is None:
            self.h_lv = lv.clone()
        else:
            h = _pad(self.h_lv, lv.shape[0], -999999.0)                                            # :213-215
            feat, _, _ = self.AFLOW(lv, h, ls)
            cat = torch.relu(self.linear(torch.cat([feat, lv], 1)))                                # :223-227
            lv = 0.0 * h + cat                                                                      # :229
            self.h_lv = lv.clone()
        ls.set_values(lv)
        return lv, ls


def make_fusion(kind, C):
    return {"linear": lambda: TemporalLinearModule(C), "maxpool": TemporalMaxPoolModule,
            "cga": lambda: CrossframeGlobalAttentionModule(C), "aflow": lambda: CrossframeLocalInterpolationModule(C),
            "lstm": lambda: LSTMModule(C), "gru": lambda: GRUModule(C)}.get(kind, lambda: None)()


class PointNetSeqModule(torch.nn.Module):
    """lattice_modules.py:343-576, experiment "none"/"slice_no_deform"/"pointnet_no_local_mean" branch"""

    def __init__(self, widths, nr_out, rnn_modules, sequence_learning):
        super().__init__()
        self.widths, self.nr_out = list(widths), nr_out
        self.layers = torch.nn.ModuleList([])
        self.kind = rnn_modules[0] if sequence_learning else "none"
        self.sequence_learning = sequence_learning
        self.fusion_module = make_fusion(self.kind, self.widths[-1] * 2) if sequence_learning else None
        self.last_conv = None

    def reset_sequence(self):
        self.fusion_module.reset_sequence()

    def forward(self, ls, distributed, indices):
        if self.last_conv is None:                                                                   # :416-440
            nr_in = distributed.shape[1] - 1
            for w in self.widths:
                lin = torch.nn.Linear(nr_in, w, bias=True)
                with torch.no_grad():
                    torch.nn.init.kaiming_normal_(lin.weight, mode="fan_in", nonlinearity="relu")
                self.layers.append(lin)
                nr_in = w
            self.last_conv = LM.ConvLatticeModule(nr_filters=self.nr_out, neighbourhood_size=1, dilation=1, bias=False)
        bary = distributed[:, -1]                                                                    # :448
        x = distributed[:, :-1]                                                                      # :452
        for i, lin in enumerate(self.layers):                                                        # :460-473
            x = lin(x)
            if i < len(self.layers) - 1:
                x = torch.relu(x)
        idx = indices.long().clone()
        idx[idx < 0] = 0                                                                             # :477-480
        red, argmax = torch_scatter.scatter_max(x, idx, dim=0)                                      # :512
        arg = argmax.clone()
        arg[argmax > argmax.shape[0]] = 0                                                            # :513-514 (Q3, literal)
        cnt = torch_scatter.scatter_add(torch.ones(idx.shape[0]), idx).unsqueeze(1)                 # :519-521
        bary_red = torch.index_select(bary, 0, arg.flatten()).view(argmax.shape[0], arg.shape[1])   # :522-523
        red = torch.cat((red, bary_red), 1)                                                          # :525
        if self.kind != "maxpool":
            red = red.masked_fill(cnt < 4, 0)                                                        # :527-530
        ls.set_values(red)
        if self.kind == "maxpool":                                                                   # :555-563
            half = red[:, : red.shape[1] // 2]
            red = red.masked_fill(half.abs().sum(1, keepdim=True) == 0, -9900)
            red, ls = self.fusion_module(red, ls)
        elif self.sequence_learning:
            red, ls = self.fusion_module(red, ls)                                                    # :565
        red = torch.index_fill(red, 0, torch.tensor([0]), 0)                                         # :569-570 (Q7)
        ls.set_values(red)
        red, ls = self.last_conv(red, ls)                                                            # :573
        ls.set_values(red)
        return red, ls


class LNNSeqOracle(torch.nn.Module):
    """seq_lattice/models.py:15-476 for the experiments without attention pooling"""

    def __init__(self, nr_classes, mp, model_cfg):
        super().__init__()
        self.sequence_learning = bool(model_cfg["sequence_learning"])
        self.rnn_modules = [k if k in _KINDS else "none" for k in (str(x).lower() for x in model_cfg["rnn_modules"])]
        assert self.rnn_modules.count("none") < len(self.rnn_modules)                               # models.py:56
        self.first_sequence = True
        self.nd = mp.nr_downsamples()
        start = mp.pointnet_start_nr_channels()
        experiment = mp.experiment()
        self.distribute = LM.DistributeLatticeModule(experiment)                                     # :62
        self.point_net_seq = PointNetSeqModule(mp.pointnet_layers(), start, self.rnn_modules, self.sequence_learning)
        widths = (start, start * 4, start * 3)                                                       # :76-152
        self.recurrent_fusion_modules = torch.nn.ModuleList(
            [make_fusion(self.rnn_modules[1 + i], widths[i]) for i in range(3)]) if self.sequence_learning else None
        self.resnet_blocks_per_down_lvl_list = torch.nn.ModuleList([])
        self.coarsens_list = torch.nn.ModuleList([])
        skips, cur = [], start
        for i in range(self.nd):                                                                     # :161-184
            blocks = torch.nn.ModuleList([])
            for _ in range(mp.nr_blocks_down_stage()[i]):
                blocks.append(LM.ResnetBlock(cur, [1, 1], [False, False], False) if i < mp.nr_levels_down_with_normal_resnet()
                              else LM.BottleneckBlock(cur, [False, False, False]))
            self.resnet_blocks_per_down_lvl_list.append(blocks)
            skips.append(cur)
            cur = int(cur * 2 * mp.compression_factor())
            self.coarsens_list.append(LM.GnReluCoarsen(cur))
        self.resnet_blocks_bottleneck = torch.nn.ModuleList(
            [LM.BottleneckBlock(cur, [False, False, False]) for _ in range(mp.nr_blocks_bottleneck())])  # :190-193
        self.finefy_list = torch.nn.ModuleList([])
        self.resnet_blocks_per_up_lvl_list = torch.nn.ModuleList([])
        for i in range(self.nd):                                                                     # :201-230
            skip = skips.pop()
            fine = int(cur / 2)
            self.finefy_list.append(LM.GnReluFinefy(fine))
            cur = skip + fine
            blocks = torch.nn.ModuleList([])
            n_up = mp.nr_blocks_up_stage()[i]
            for j in range(n_up):
                last = (j == n_up - 1) and (i == self.nd - 1)
                blocks.append(LM.ResnetBlock(cur, [1, 1], [False, last], False) if i >= self.nd - mp.nr_levels_up_with_normal_resnet()
                              else LM.BottleneckBlock(cur, [False, False, last]))
            self.resnet_blocks_per_up_lvl_list.append(blocks)
        self.slice_fast_cuda = LM.SliceFastCUDALatticeModule(nr_classes=nr_classes, dropout_prob=mp.dropout_last_layer(),
                                                             experiment=experiment)                  # :232

    def reset_sequence(self):                                                                        # :236-243
        self.first_sequence = True
        if self.sequence_learning:
            self.point_net_seq.reset_sequence()
            for m in self.recurrent_fusion_modules:
                if m is not None:
                    m.reset_sequence()

    def _fuse(self, k, lv, ls):
        m = self.recurrent_fusion_modules[k] if self.sequence_learning else None
        return m(lv, ls) if m is not None else (lv, ls)

    def forward(self, ls, positions, values, early_return=False):
        seq, rnn = self.sequence_learning, self.rnn_modules
        reset = not (seq and not self.first_sequence)                                                # :287-289
        ls, distributed, indices, weights = self.distribute(ls, positions, values, reset)            # :298
        lv, ls = self.point_net_seq(ls, distributed, indices)                                        # :303
        if early_return and seq and rnn[1:] == ["none"] * 3:                                         # :307-309
            self.first_sequence = False
            return lv, lv, ls
        saved = []
        for i in range(self.nd):                                                                     # :314-353
            for b in self.resnet_blocks_per_down_lvl_list[i]:
                lv, ls = b(lv, ls)
            saved.append((ls, lv))
            if i == 0:
                lv, ls = self._fuse(0, lv, ls)
                if early_return and seq and rnn[2:] == ["none"] * 2:
                    self.first_sequence = False
                    return lv, lv, ls
            lv, ls = self.coarsens_list[i](lv, ls)
        for b in self.resnet_blocks_bottleneck:                                                      # :361-363
            lv, ls = b(lv, ls)
        lv, ls = self._fuse(1, lv, ls)                                                               # :381-382
        for i in range(self.nd):                                                                     # :390-430
            fine_ls, fine_lv = saved.pop()
            lv, ls = self.finefy_list[i](lv, ls, fine_ls)
            lv = torch.cat((lv, fine_lv), 1)
            if i == self.nd - 1:
                lv, ls = self._fuse(2, lv, ls)
                if early_return and seq:
                    self.first_sequence = False
                    return lv, lv, ls
        for b in self.resnet_blocks_per_up_lvl_list[self.nd - 1]:                                    # :435-437 (Q1)
            lv, ls = b(lv, ls)
        sv = self.slice_fast_cuda(lv, ls, positions, indices, weights)                               # :465
        self.first_sequence = False
        return torch.log_softmax(sv, 1), sv, ls                                                      # :474-476


class OracleWindowRunner:
    """The window loop of test_ln.py:149-231 over the oracle model, CPU only."""

    def __init__(self, cfg_path, nr_classes=26):
        self.cfg_path = cfg_path
        with open(cfg_path) as f:
            model_cfg = hjson.loads(f.read())["model"]
        self.model = LNNSeqOracle(nr_classes, ModelParams.create(cfg_path), model_cfg)
        self.model.train(False)
        self.lattice = None

    def materialise_parameters(self, frames, state_dict_fn=None):
        self.infer_window(frames)
        if state_dict_fn is None:
            from temporal_latticenet_b200.seeding import seeded_state as state_dict_fn
        shapes = {k: tuple(v.shape) for k, v in self.model.state_dict().items()}
        self.model.load_state_dict(state_dict_fn(shapes))
        return self

    def infer_window(self, frames, collect=None):
        """frames: list of (positions, values) numpy arrays.  Returns log-softmax [N, classes] of the
        last frame; `collect` (a list) receives (first output, vertex count) per frame."""
        self.model.reset_sequence()
        ls = Lattice.create(self.cfg_path, "lattice")
        out = None
        with torch.no_grad():
            for i, (p, v) in enumerate(frames):
                out, raw, ls = self.model(ls, torch.from_numpy(p), torch.from_numpy(v), i != len(frames) - 1)
                if collect is not None:
                    collect.append((out, raw, ls.nr_lattice_vertices()))
        self.lattice = ls
        return out
