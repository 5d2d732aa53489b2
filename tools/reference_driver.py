#!/usr/bin/env python3
"""Runs the REFERENCE's own, unmodified model code -- seq_lattice/models.py::LNN_SEQ and seq_lattice/lattice_modules.py
-- on the GPU over temporal_latticenet_b200/shims (the drop-in boundary of north_star: `latticenet`, `latticenet_py`,
`torch_scatter`, `termcolor` backed by the sm_100a kernels).

The two reference files are NOT part of this repository.  `__graft_entry__.build()` copies them byte for byte from
/root/reference into baseline/_ref/ (git-ignored; it travels to the GPU box like a built .so) and this script verifies
their SHA-256 against the values recorded in tests/golden/reference_files.json before importing them.

  python tools/reference_driver.py golden <name> <out.npz>     one golden window (tests/golden/<name>.npz inputs) -> outputs
  python tools/reference_driver.py bench <steps> <warmup>      scans/s of the reference-owned Python driving our kernels

Always a separate process: the oracle's CPU shims use the same module names.
"""
import hashlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
SHIMS = os.path.join(REPO, "temporal_latticenet_b200", "shims")
FILES = ("seq_lattice/models.py", "seq_lattice/lattice_modules.py", "cfgParser.py")
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def fixture_status():
    """None when the fixture is present and unmodified, else a reason"""
    with open(os.path.join(REPO, "tests", "golden", "reference_files.json")) as f:
        want = json.load(f)
    for rel in FILES:
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            return "baseline/_ref/%s is missing (run __graft_entry__.build() where /root/reference is mounted)" % rel
        with open(path, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != want[rel]:
                return "baseline/_ref/%s differs from the reference file it claims to be" % rel
    return None


def _import_reference():
    why = fixture_status()
    if why:
        raise RuntimeError(why)
    sys.path.insert(0, REF)
    sys.path.insert(0, SHIMS)
    import latticenet
    assert os.path.abspath(latticenet.__file__).startswith(SHIMS)
    from cfgParser import cfgParser                # reference file
    from seq_lattice.models import LNN_SEQ         # reference file
    import seq_lattice.models as M
    assert os.path.abspath(M.__file__).startswith(REF)
    return LNN_SEQ, cfgParser, latticenet.Lattice, latticenet.ModelParams


def _window(model, Lattice, cfg, frames, collect=None):
    """the reference's frame loop (test_ln.py:149-166)"""
    import torch
    lattice = Lattice.create(cfg, "lattice")
    out = None
    for i, (p, v) in enumerate(frames):
        early = i != len(frames) - 1
        out, raw, lattice = model(lattice, p, v, early, with_gradient=False)
        if collect is not None:
            collect.append((out.detach().cpu().numpy().copy(), raw.detach().cpu().numpy().copy(), int(lattice.nr_lattice_vertices())))
    return out, lattice


def _build(cfg, nr_classes, frames):
    """test_ln.py:165-185: first window creates the lazy parameters, load the (seeded) state-dict, reset"""
    import torch
    from temporal_latticenet_b200.seeding import seeded_state
    LNN_SEQ, cfgParser, Lattice, ModelParams = _import_reference()
    model = LNN_SEQ(nr_classes, ModelParams.create(cfg), cfgParser(cfg)).to("cuda")
    model.train(False)
    with torch.no_grad():
        _window(model, Lattice, cfg, frames)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(seeded_state(shapes))
        model.reset_sequence()
    return model, Lattice, shapes


def golden(name, out_path):
    import hjson
    import numpy as np
    import torch
    gdir = os.path.join(REPO, "tests", "golden")
    z = np.load(os.path.join(gdir, name + ".npz"))
    with open(os.path.join(gdir, name + ".json")) as f:
        meta = json.load(f)
    with open(os.path.join(REPO, "configs", "lnn_eval_semantic_kitti.cfg")) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = meta["rnn_modules"]
    cfg["model"]["sequence_learning"] = meta["sequence_learning"]
    cfg["loader_semantic_kitti"]["frames_per_seq"] = meta["frames"]
    cfg_path = out_path + ".cfg"
    with open(cfg_path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")
    frames = [(torch.from_numpy(z["pos%d" % i]).to(dev), torch.from_numpy(z["val%d" % i]).to(dev)) for i in range(meta["frames"])]
    model, Lattice, shapes = _build(cfg_path, meta["nr_classes"], frames)
    collect = []
    with torch.no_grad():
        _, lattice = _window(model, Lattice, cfg_path, frames, collect)
    arrays = {"logits": collect[-1][1], "keys0": lattice.hash_table.keys().cpu().numpy()}
    for i, (a, b, nv) in enumerate(collect):
        arrays["nv%d" % i] = np.int64(nv)
        if i < len(collect) - 1:
            arrays["out%d" % i] = a
    np.savez(out_path, **arrays)
    with open(out_path + ".shapes.json", "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f)
    os.remove(cfg_path)


def bench(steps, warmup):
    """scans/s of config 3 with the reference's Python (PointNetSeqModule's three nn.Linear + scatter_max, nn.GRUCell,
    the unfused AFlow) driving the kernels eagerly: what a user gets by ONLY swapping the imports."""
    import torch
    from temporal_latticenet_b200 import synthetic
    dev = torch.device("cuda:0")
    cfg = os.path.join(REPO, "configs", "lnn_eval_semantic_kitti.cfg")
    windows = [synthetic.window(1000 + i, frames=4, scope=3) for i in range(2)]
    devw = [[(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w] for w in windows]
    model, Lattice, _ = _build(cfg, 26, devw[0])
    with torch.no_grad():
        for i in range(max(warmup, 1)):
            model.reset_sequence()
            _window(model, Lattice, cfg, devw[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            model.reset_sequence()
            out, _ = _window(model, Lattice, cfg, devw[i % 2])
            labels = out.argmax(1)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    print(json.dumps({"driver": "reference seq_lattice/models.py + lattice_modules.py over temporal_latticenet_b200/shims (eager)",
                      "scans_per_s": 4 * steps / wall, "ms_per_window": 1e3 * wall / steps, "device_ms_per_window": e0.elapsed_time(e1) / steps,
                      "steps": steps, "labels": int(labels.shape[0])}))


if __name__ == "__main__":
    if sys.argv[1] == "golden":
        golden(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "bench":
        bench(int(sys.argv[2]), int(sys.argv[3]))
    else:
        raise SystemExit(__doc__)
