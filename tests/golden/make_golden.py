"""Generates tests/golden/*.npz by running the REFERENCE's model code on the CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

What runs: /root/reference/seq_lattice/models.py::LNN_SEQ and seq_lattice/lattice_modules.py,
unmodified, imported from where they lie, over the oracle shims in oracle/shims (CPU restatement of
the absent `latticenet`, `latticenet_py`, `torch_scatter` dependencies).  The reference hard-codes
`.to("cuda")` / torch.cuda.FloatTensor (lattice_modules.py:109,291,309-311,424,505,519,569); a
context manager maps those to the CPU for the duration of the run -- no reference file is edited.

Protocol = the reference's own checkpoint flow (test_ln.py:165-185): one full window creates the
lazy parameters, the state-dict is loaded, reset_sequence(), fresh Lattice, the window is re-run.
The checkpoint itself is missing (.MISSING_LARGE_BLOBS:1) so the state-dict is seeded by name.
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REFERENCE = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "shims"))
sys.path.insert(0, REFERENCE)

from tests.helpers import CFG, seeded_state, small_window  # noqa: E402


@contextlib.contextmanager
def cpu_as_cuda():
    t_to, m_to = torch.Tensor.to, torch.nn.Module.to
    had = hasattr(torch.cuda, "FloatTensor")
    old_ft = getattr(torch.cuda, "FloatTensor", None)

    def fix(args, kwargs):
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
        if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
            kwargs = dict(kwargs, device="cpu")
        return args, kwargs

    def tensor_to(self, *a, **k):
        a, k = fix(a, k)
        return t_to(self, *a, **k)

    def module_to(self, *a, **k):
        a, k = fix(a, k)
        return m_to(self, *a, **k)

    torch.Tensor.to, torch.nn.Module.to = tensor_to, module_to
    torch.cuda.FloatTensor = torch.FloatTensor
    try:
        yield
    finally:
        torch.Tensor.to, torch.nn.Module.to = t_to, m_to
        if had:
            torch.cuda.FloatTensor = old_ft


def run_window(model, Lattice, cfg, frames, collect):
    lattice = Lattice.create(cfg, "lattice")
    outs = []
    for i, (p, v) in enumerate(frames):
        early = i != len(frames) - 1
        a, b, lattice = model(lattice, torch.from_numpy(p), torch.from_numpy(v), early, with_gradient=False)
        if collect:
            outs.append((a.detach().numpy().copy(), b.detach().numpy().copy(), lattice.nr_lattice_vertices()))
    return outs, lattice


def make(name, rnn_modules, seed, nr_classes=26, sequence_learning=True, frames_n=4, radius=9.0, max_points=6000, keep_frames=True):
    import hjson
    from cfgParser import cfgParser  # reference file, unmodified
    from latticenet import Lattice, ModelParams
    from seq_lattice.models import LNN_SEQ  # reference file, unmodified

    # own cfg, same keys; rnn_modules / sequence_learning varied per golden
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = rnn_modules
    cfg["model"]["sequence_learning"] = sequence_learning
    cfg["loader_semantic_kitti"]["frames_per_seq"] = frames_n
    tmp_cfg = os.path.join(HERE, "_tmp_%s.cfg" % name)
    with open(tmp_cfg, "w") as f:
        f.write(hjson.dumps(cfg))
    try:
        frames = small_window(seed=seed, frames=frames_n, radius=radius, max_points=max_points)
        with cpu_as_cuda(), torch.no_grad():
            torch.manual_seed(0)
            model = LNN_SEQ(nr_classes, ModelParams.create(tmp_cfg), cfgParser(tmp_cfg)).to("cuda")
            model.train(False)
            run_window(model, Lattice, tmp_cfg, frames, collect=False)  # creates the lazy parameters
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            model.load_state_dict(seeded_state(shapes))
            model.reset_sequence()
            outs, lattice = run_window(model, Lattice, tmp_cfg, frames, collect=True)
    finally:
        os.remove(tmp_cfg)
    arrays = {}
    for i, (p, v) in enumerate(frames):
        arrays["pos%d" % i], arrays["val%d" % i] = p, v
    for i, (a, b, nv) in enumerate(outs):
        if keep_frames and i < len(outs) - 1:
            arrays["out%d" % i] = a  # lv after late fusion on the early-return frames (models.py:427-430)
        arrays["nv%d" % i] = np.int64(nv)
    arrays["logits"] = outs[-1][1]
    arrays["keys0"] = lattice.hash_table.keys()
    meta = {"name": name, "rnn_modules": rnn_modules, "seed": seed, "nr_classes": nr_classes,
            "sequence_learning": sequence_learning, "frames": frames_n, "radius": radius, "max_points": max_points,
            "shapes": {k: list(v) for k, v in shapes.items()}}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(name, "points", [f[0].shape[0] for f in frames], "V0 per frame", [o[2] for o in outs],
          "logits", outs[-1][1].shape, "absmax", float(np.abs(outs[-1][1]).max()))


if __name__ == "__main__":
    only = sys.argv[1:]
    _make = make

    def make(name, *a, **k):  # noqa: F811
        if not only or name in only:
            _make(name, *a, **k)

    make("gru_gru_aflow_gru", ["gru", "gru", "aflow", "gru"], seed=1)
    small = dict(frames_n=3, radius=6.0, max_points=2500, keep_frames=False)
    make("lstm_cga_linear_maxpool", ["lstm", "cga", "linear", "maxpool"], seed=2, **small)
    make("maxpool_aflow_lstm_cga", ["maxpool", "aflow", "lstm", "cga"], seed=3, **small)
    make("aflow_x4", ["aflow", "aflow", "aflow", "aflow"], seed=4, **small)
    make("linear_none_none_gru", ["linear", "none", "none", "gru"], seed=5, **small)
    # BASELINE config 2 (single-frame LatticeNet, 20 classes).  With sequence_learning:false the
    # reference itself raises AttributeError at models.py:424 (recurrent_fusion_modules is only
    # created under sequence_learning, models.py:74,155), so the single-frame golden is a one-frame
    # window with sequence_learning:true: every fusion module takes its t==0 identity branch.
    make("single_frame", ["gru", "gru", "aflow", "gru"], seed=6, nr_classes=20, sequence_learning=True,
         frames_n=1, radius=8.0, max_points=5000, keep_frames=False)
