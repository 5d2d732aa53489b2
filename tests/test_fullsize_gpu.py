"""Parity at BASELINE.json's FULL size (one synthetic SemanticKITTI-shaped window: 4 scans of ~125k points).

Lattice structure of all three levels, barycentric weights and neighbour tables: bit-exact against the C oracle.
Feature path: the whole timed window (and a config-2 single frame, a config-5 accumulated cloud) against
oracle/window_oracle.py, which runs a full-size window in seconds on the host (second half of this file); plus
size-independent properties (partition of unity of splat/slice, probabilities summing to one, append-only vertex
ids) and the three execution modes (eager, CUDA-graph replay, several windows in flight) agreeing with each other."""
import numpy as np
import pytest
import torch

from tests.helpers import CFG, canonical_order

pytestmark = pytest.mark.gpu

from oracle import lattice_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def window():
    from temporal_latticenet_b200 import synthetic
    return synthetic.window(1000, frames=4, scope=3)


def test_full_size_lattice_structure_is_bit_exact(window):
    from temporal_latticenet_b200.lattice import Lattice
    dev = torch.device("cuda:0")
    ls = Lattice(100000, 0.6, device=dev)
    t0, t1, t2 = O.OracleTable(100000), O.OracleTable(100000), O.OracleTable(100000)
    prev = 0
    for f, (p, v) in enumerate(window):
        rows, idx, w = ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), f == 0)
        c1 = ls.create_coarse_verts()
        c2 = c1.create_coarse_verts()
        _, o_idx, o_w = t0.distribute(p, v, O.scale_factors(0.6))
        t1.insert_points(p, O.scale_factors(1.2))
        t2.insert_points(p, O.scale_factors(2.4))
        assert np.array_equal(idx.cpu().numpy(), o_idx)
        assert np.array_equal(w.cpu().numpy(), o_w)
        for l, t in ((ls, t0), (c1, t1), (c2, t2)):
            k = l.hash_table.keys().cpu().numpy()
            assert np.array_equal(k, t.keys())                                      # ids in insertion order
            assert np.array_equal(k[canonical_order(k)], t.keys()[canonical_order(t.keys())])
            assert len({tuple(r) for r in k}) == k.shape[0]                         # no duplicate vertex
        assert ls.nr_lattice_vertices() >= prev                                      # append-only across frames
        prev = ls.nr_lattice_vertices()
    assert np.array_equal(ls.neighbours().cpu().numpy(), t0.neighbours())
    assert np.array_equal(c1.neighbours(ls, mode=1).cpu().numpy(), t1.neighbours(t0, mode=1))
    assert np.array_equal(ls.neighbours(c1, mode=2).cpu().numpy(), t0.neighbours(t1, mode=2))
    assert prev < 100000 and ls.hash_table.nr_overflowed() == 0


def test_full_size_splat_slice_properties(window):
    from temporal_latticenet_b200 import funcs
    from temporal_latticenet_b200.lattice import Lattice
    dev = torch.device("cuda:0")
    p, v = window[0]
    ls = Lattice(100000, 0.6, device=dev)
    pt = torch.from_numpy(p).to(dev)
    lv, idx, w = funcs.SplatLattice.apply(ls, pt, torch.from_numpy(v).to(dev))
    V = ls.nr_lattice_vertices()
    assert abs(float(lv[:, -1].double().sum()) - p.shape[0]) < 1e-3 * p.shape[0] ** 0.5   # mass = number of points
    assert abs(float(lv[:, 0].double().sum()) - float(v.astype(np.float64).sum())) < 1e-2  # linearity in the values
    ones = funcs.SliceLattice.apply(torch.ones(V, 4, device=dev), ls, pt, idx, w)
    assert float((ones - 1).abs().max()) < 1e-5                                             # partition of unity
    tab = O.OracleTable(100000)
    _, o_idx, o_w = tab.distribute(p, np.zeros((p.shape[0], 1), np.float32), O.scale_factors(0.6))
    want = O.splat(v, o_idx, o_w, V)
    np.testing.assert_allclose(lv.cpu().numpy(), want, rtol=2e-4, atol=2e-4)


def test_full_size_window_modes_agree(window):
    from temporal_latticenet_b200.engine import MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in window]
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(fd, seeded_state)
    want = eager.infer_window_device(fd)
    counts = []
    lvl = eager.lattice
    while lvl is not None:
        counts.append(lvl.nr_lattice_vertices())
        lvl = lvl._coarse
    assert counts[0] > counts[1] > counts[2] > 0
    prob = want.exp().sum(1)
    assert float((prob - 1).abs().max()) < 1e-4 and torch.isfinite(want).all()
    again = eager.infer_window_device(fd)
    assert float((again - want).abs().max()) < 1e-4 * float(want.abs().max())   # run-to-run: only atomics order differs
    multi = MultiWindowRunner(CFG, 26, dev, lanes=2).prepare(fd, seeded_state)
    outs = multi.infer_windows_device([fd, fd])
    torch.cuda.synchronize()
    assert multi.counts_ok()
    for o in outs:
        assert o.shape == want.shape
        assert float((o - want).abs().max()) < 1e-4 * float(want.abs().max())
        assert float((o.argmax(1) == want.argmax(1)).float().mean()) > 0.999


# ---------------------------------------------------------------------------------------------------------------
# Feature-level parity at FULL size against the CPU oracle (oracle/window_oracle.py runs a whole 4 x 125k-point
# window in about a second): the window bench.py times (seed 1000), a config-2 single frame (20 classes) and a
# config-5 window ([aflow x4] on the 4 scans accumulated into one ~500k-point cloud).
#
# Tolerances (fp32, stated per north_star): every compared tensor within FULL_TOL = 1e-4 of its absolute maximum
# at the 99.9th percentile of its elements; the class decision equal wherever the oracle's top-2 margin exceeds
# 4 x FULL_TOL x absmax.  The percentile (not the maximum) is asserted because PointNet's scatter_max is followed by
# a gather of the barycentric weight OF THE ARG-MAX ROW (lattice_modules.py:512-525): where two rows of a vertex tie
# to within fp32 rounding the winner -- and with it a whole input channel of that vertex -- legitimately differs
# between two correct fp32 implementations.  The maximum is held to OUTLIER_TOL and the share of elements beyond
# FULL_TOL is printed.
# ---------------------------------------------------------------------------------------------------------------
FULL_TOL = 1e-4
OUTLIER_TOL = 5e-2


def _cmp(name, got, want, tol=FULL_TOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    fin = np.isfinite(want)
    if not fin.all():
        # quirk Q5 (AFlow's 0/0 on vertices whose whole neighbourhood is zero rows, lattice_modules.py:321): the reference
        # itself produces NaN there and GroupNorm spreads it; parity then means producing NaN in the same places
        same = float((np.isfinite(got) == fin).mean())
        print("%-28s oracle non-finite share %.4f, finite-mask agreement %.6f" % (name, 1.0 - float(fin.mean()), same))
        assert same > 0.999, (name, same)
        if not fin.any():
            return None
        got, want = got[fin & np.isfinite(got)], want[fin & np.isfinite(got)]
    scale = float(np.abs(want).max()) + 1e-30
    err = np.abs(got - want) / scale
    q999 = float(np.quantile(err, 0.999))
    beyond = float((err > tol).mean())
    print("%-28s shape %-16s absmax %.3e  err/absmax: median %.2e  p99.9 %.2e  max %.2e  beyond tol %.2e"
          % (name, got.shape, scale, float(np.median(err)), q999, float(err.max()), beyond))
    assert q999 < tol, (name, q999)
    assert float(err.max()) < OUTLIER_TOL, (name, float(err.max()))
    return err


def _decisions(name, got, want, tol=FULL_TOL):
    ok = np.isfinite(want).all(1) & np.isfinite(got).all(1)
    if not ok.any():
        return
    got, want = got[ok], want[ok]
    top2 = np.sort(want, 1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4 * tol * np.abs(want).max()
    agree = float((got.argmax(1)[clear] == want.argmax(1)[clear]).mean())
    print("%-28s clear-margin points %.4f of all, label agreement on them %.6f, overall %.6f"
          % (name, float(clear.mean()), agree, float((got.argmax(1) == want.argmax(1)).mean())))
    assert agree > 0.9995, (name, agree)


def _oracle_window(cfg, nr_classes, frames):
    from oracle import window_oracle as WO
    orc = WO.OracleWindowRunner(cfg, nr_classes).materialise_parameters(frames[:2])  # two frames: the t > 0 branches create AFlow's lazy parameters too
    collect = []
    out = orc.infer_window(frames, collect=collect)
    counts, lvl = [], orc.lattice
    while lvl is not None:
        counts.append(lvl.nr_lattice_vertices())
        lvl = getattr(lvl, "_coarse", None)
    return out.numpy(), [(a.numpy(), b.numpy(), n) for a, b, n in collect], counts


def _eager_frames(runner, frames_dev):
    """the window loop of WindowRunner.infer_window_device, keeping every frame's outputs"""
    from temporal_latticenet_b200 import ops
    outs = []
    ls = runner.new_lattice()
    last = len(frames_dev) - 1
    with torch.no_grad(), ops.tc_operands(runner.operands, runner.range_flag):
        for i, (p, v) in enumerate(frames_dev):
            a, b, ls = runner.model(ls, p, v, i != last, False)
            outs.append((a.clone(), b.clone(), ls.nr_lattice_vertices()))
    runner.lattice = ls
    assert not runner.range_raised()
    return outs


def test_full_size_timed_window_matches_oracle(window):
    """BASELINE config 3 on the very window bench.py times: per-frame late-fusion features (the three
    early-return tensors), last-frame logits and log-softmax, eager runner AND the multi-window graph runner."""
    from temporal_latticenet_b200.engine import MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    want, per_frame, counts = _oracle_window(CFG, 26, window)
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in window]
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(fd, seeded_state)
    outs = _eager_frames(eager, fd)
    got_counts, lvl = [], eager.lattice
    while lvl is not None:
        got_counts.append(lvl.nr_lattice_vertices())
        lvl = lvl._coarse
    assert got_counts == counts[: len(got_counts)], (got_counts, counts)
    for t in range(len(window)):
        assert outs[t][2] == per_frame[t][2], "vertex count after frame %d" % t
    for t in range(len(window) - 1):
        _cmp("late-fusion lv, frame %d" % t, outs[t][0].cpu().numpy(), per_frame[t][0])
    _cmp("logits (eager)", outs[-1][1].cpu().numpy(), per_frame[-1][1])
    _cmp("log-softmax (eager)", outs[-1][0].cpu().numpy(), want)
    _decisions("labels (eager)", outs[-1][0].cpu().numpy(), want)
    multi = MultiWindowRunner(CFG, 26, dev, lanes=2).prepare(fd, seeded_state)
    got = [o.cpu().numpy() for o in multi.infer_windows_device([fd, fd])]
    torch.cuda.synchronize()
    assert multi.counts_ok()
    for i, g in enumerate(got):
        _cmp("log-softmax (graph lane %d)" % i, g, want)
        _decisions("labels (graph lane %d)" % i, g, want)


def test_full_size_single_frame_matches_oracle(window, tmp_path):
    """BASELINE config 2: one full scan through the LatticeNet forward (conv / coarsen / finefy / slice_classify), 20 classes"""
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    frames = window[:1]
    want, per_frame, counts = _oracle_window(CFG, 20, frames)
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in frames]
    eager = WindowRunner(CFG, 20, dev).materialise_parameters(fd, seeded_state)
    outs = _eager_frames(eager, fd)
    assert outs[0][2] == per_frame[0][2]
    _cmp("config2 logits", outs[0][1].cpu().numpy(), per_frame[0][1])
    _cmp("config2 log-softmax", outs[0][0].cpu().numpy(), want)
    _decisions("config2 labels", outs[0][0].cpu().numpy(), want)


def test_full_size_accumulated_aflow_matches_oracle(window, tmp_path):
    """BASELINE config 5: rnn_modules [aflow x4]; (a) the 4 scans accumulated into ONE ~500k-point cloud
    (accumulate_clouds: kitti_dataloader.py:198-201 -- a single "frame", every AFlow takes its t == 0 branch),
    (b) the same modules over the 4-frame window (the t > 0 AFlow branches at full size)."""
    import hjson
    import os
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = ["aflow", "aflow", "aflow", "aflow"]
    cfg["lattice_gpu"]["hash_table_capacity"] = 200000
    path = os.path.join(str(tmp_path), "aflow_x4.cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")
    acc = [(np.ascontiguousarray(np.concatenate([p for p, _ in window], 0)),
            np.ascontiguousarray(np.concatenate([v for _, v in window], 0)))]
    for name, frames in (("config5 accumulated", acc), ("config5 4-frame", window)):
        want, per_frame, counts = _oracle_window(path, 26, frames)
        fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in frames]
        eager = WindowRunner(path, 26, dev).materialise_parameters(fd, seeded_state)
        outs = _eager_frames(eager, fd)
        for t in range(len(frames)):
            assert outs[t][2] == per_frame[t][2], (name, t)
        _cmp(name + " logits", outs[-1][1].cpu().numpy(), per_frame[-1][1])
        _cmp(name + " log-softmax", outs[-1][0].cpu().numpy(), want)
        _decisions(name + " labels", outs[-1][0].cpu().numpy(), want)
