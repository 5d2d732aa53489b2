"""Graph-replay runner (static capacities, device-side sizes, one CUDA graph per frame kind) against the
eager runner on the same windows: same vertex counts, same log-softmax within 1e-4 of abs-max (the two
paths run the same kernels; only the order of the floating-point atomics of the GroupNorm statistics
differs), same predicted labels."""
import numpy as np
import pytest
import torch

from tests.helpers import CFG, small_window

pytestmark = pytest.mark.gpu


def _window(seed, frames=4, n=6000):
    return small_window(seed=seed, frames=frames, radius=9.0, max_points=n)


def test_graph_runner_matches_eager_runner():
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    wins = [_window(1), _window(2, n=5500), _window(3, n=5000)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    graph = GraphWindowRunner(CFG, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    assert graph.supported
    graph.plan(to_dev(wins[0]))
    graph.capture(to_dev(wins[0]))
    assert len(graph.graphs) == 3  # first / middle / last
    for rep in range(2):  # replaying twice checks that no state leaks between windows
        for w in wins:
            fd = to_dev(w)
            want = eager.infer_window_device(fd).cpu().numpy()
            got = graph.infer_window_device(fd).cpu().numpy()
            assert graph.counts_ok()
            counts = []
            lvl = graph.static_lattice
            while lvl is not None:
                counts.append(int(lvl.hash_table.count_tensor().cpu()))
                lvl = lvl._coarse
            ecounts = []
            lvl = eager.lattice
            while lvl is not None:
                ecounts.append(lvl.nr_lattice_vertices())
                lvl = lvl._coarse
            assert counts == ecounts
            assert got.shape == want.shape
            err = float(np.abs(got - want).max()) / float(np.abs(want).max())
            assert err < 1e-4, err
            assert (got.argmax(1) == want.argmax(1)).mean() > 0.999
    # end-to-end entry point with pinned host buffers
    host = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in wins[1]]
    labels = graph.infer_window(host).clone()
    want = eager.infer_window(host)
    assert (labels == want).float().mean() > 0.999 and graph.fallbacks == 0


def test_graph_runner_falls_back_when_capacity_is_exceeded():
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    small = small_window(seed=4, frames=4, radius=5.0, max_points=2500)
    big = small_window(seed=5, frames=4, radius=12.0, max_points=2500)   # same point count, far more vertices
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(to_dev(small), seeded_state)
    graph = GraphWindowRunner(CFG, 26, dev, headroom=1.05).materialise_parameters(to_dev(small), seeded_state)
    graph.plan(to_dev(small))
    graph.capture(to_dev(small))
    host = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in big]
    labels = graph.infer_window(host).clone()
    assert graph.fallbacks == 1
    assert torch.equal(labels, eager.infer_window(host))


def test_multi_window_runner_matches_eager_runner():
    """three windows in flight on three streams give the results of running them one after the other"""
    from temporal_latticenet_b200.engine import MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    wins = [_window(6), _window(7, n=5200), _window(8, n=4800)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    devw = [to_dev(w) for w in wins]
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(devw[0], seeded_state)
    multi = MultiWindowRunner(CFG, 26, dev, lanes=3).prepare(devw[0], seeded_state, devw)
    want = [eager.infer_window_device(w).cpu().numpy() for w in devw]
    for rep in range(3):
        order = devw[rep:] + devw[:rep]          # rotate which lane gets which window
        outs = multi.infer_windows_device(order)
        torch.cuda.synchronize()
        assert multi.counts_ok()
        for o, w in zip(outs, (want[rep:] + want[:rep])):
            got = o.cpu().numpy()
            assert got.shape == w.shape
            assert float(np.abs(got - w).max()) / float(np.abs(w).max()) < 1e-4
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in wins]
    labels = [l.clone() for l in multi.infer_windows(host)]
    for l, w in zip(labels, want):
        assert (l.numpy() == w.argmax(1)).mean() > 0.999


@pytest.mark.parametrize("rnn", [["aflow", "aflow", "aflow", "aflow"], ["lstm", "none", "none", "gru"], ["gru", "gru", "aflow", "gru"],
                                 ["maxpool", "cga", "linear", "gru"], ["linear", "maxpool", "cga", "lstm"], ["cga", "linear", "maxpool", "maxpool"]])
def test_graph_runner_other_fusion_configs(rnn, tmp_path):
    """static-capacity graphs for every fusion kind that has a device-side-size implementation (BASELINE config 5
    uses aflow x4); 3-frame windows exercise first / middle / last graphs"""
    import hjson
    import os
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = rnn
    path = os.path.join(str(tmp_path), "cfg_%s.cfg" % "_".join(rnn))
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")
    wins = [_window(11, frames=3, n=4000), _window(12, frames=3, n=3600)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(path, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    graph = GraphWindowRunner(path, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    assert graph.supported
    graph.plan(to_dev(wins[0]))
    graph.capture(to_dev(wins[0]))
    for rep in range(2):
        for w in wins:
            fd = to_dev(w)
            want = eager.infer_window_device(fd).cpu().numpy()
            got = graph.infer_window_device(fd).cpu().numpy()
            assert graph.counts_ok()
            finite = np.isfinite(want)
            assert np.array_equal(finite, np.isfinite(got))   # AFlow's 0/0 rows (quirk Q5) stay where the eager path has them
            if finite.any():   # (an early AFlow's NaN reaches every vertex through the GroupNorm statistics, as in the reference)
                err = float(np.abs(got[finite] - want[finite]).max()) / float(np.abs(want[finite]).max())
                assert err < 1e-4, err


def test_every_fusion_kind_has_a_static_capacity_path(tmp_path):
    """linear / maxpool / cga (lattice_modules.py:70-185) run inside the captured graphs too (round 1 sent them to the eager runner)"""
    import hjson
    import os
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = ["maxpool", "cga", "linear", "gru"]
    path = os.path.join(str(tmp_path), "cfg_all_kinds.cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")
    w = _window(13, frames=3, n=3000)
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]
    g = GraphWindowRunner(path, 26, dev).materialise_parameters(fd, seeded_state)
    assert g.supported
    g.capture(fd)
    out = g.infer_window_device(fd)
    assert out.shape == (w[-1][0].shape[0], 26) and len(g.graphs) == 3 and g.counts_ok()


def test_more_points_than_planned_runs_eagerly():
    from temporal_latticenet_b200.engine import MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    small, big = _window(21, n=2000), _window(22, n=9000)
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(to_dev(small), seeded_state)
    multi = MultiWindowRunner(CFG, 26, dev, lanes=2).prepare(to_dev(small), seeded_state)
    assert big[0][0].shape[0] > multi.lanes[0].caps["n"]
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in (big, small)]
    labels = [l.clone() for l in multi.infer_windows(host)]
    assert multi.lanes[0].fallbacks == 1 and multi.lanes[1].fallbacks == 0
    for l, w in zip(labels, (big, small)):
        want = eager.infer_window_device(to_dev(w)).argmax(1).cpu()
        assert (l == want).float().mean() > 0.999
    out = multi.lanes[0].infer_window_device(to_dev(big))
    assert out.shape[0] == big[-1][0].shape[0] and multi.lanes[0].fallbacks == 2


def test_graph_runner_single_accumulated_frame(tmp_path):
    """BASELINE config 5 shape: rnn_modules aflow x4 with accumulate_clouds -- the window's scans are concatenated into
    ONE frame (kitti_dataloader.py:198-201), so every fusion module only takes its t == 0 branch and the engine has a
    single frame kind (first AND last)."""
    import hjson
    import os
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = ["aflow", "aflow", "aflow", "aflow"]
    cfg["loader_semantic_kitti"]["accumulate_clouds"] = True
    path = os.path.join(str(tmp_path), "cfg_accum.cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")

    def accumulated(seed):
        w = _window(seed, frames=4, n=3000)
        return [(np.concatenate([p for p, _ in w], 0), np.concatenate([v for _, v in w], 0))]
    wins = [accumulated(31), accumulated(32)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(path, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    graph = GraphWindowRunner(path, 26, dev).materialise_parameters(to_dev(wins[0]), seeded_state)
    graph.plan(to_dev(wins[0]))
    graph.capture(to_dev(wins[0]))
    assert list(graph.graphs) == [(True, True)]
    for w in wins + wins:
        fd = to_dev(w)
        want = eager.infer_window_device(fd).cpu().numpy()
        got = graph.infer_window_device(fd).cpu().numpy()
        assert graph.counts_ok() and got.shape == want.shape == (w[0][0].shape[0], 26)
        assert float(np.abs(got - want).max()) / float(np.abs(want).max()) < 1e-4


def test_fp16_range_flag_reruns_the_window_with_tf32_operands():
    """The fp16 hi/lo operand form is only a fast path: when an activation leaves the fp16 range the kernels raise the
    device-side flag, and every runner redoes the window with the tf32 operands.  Reflectance values of 1e7 push the
    PointNet activations far beyond 65504 / 2^5; the results must equal those of a runner that uses tf32 from the start,
    on the eager path, on the graph path and through submit / collect."""
    from temporal_latticenet_b200.engine import GraphWindowRunner, MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    normal = _window(31, n=4000)
    huge = [(p, (v * 1e7).astype(np.float32)) for p, v in _window(32, n=4000)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    safe = WindowRunner(CFG, 26, dev, operands="tf32").materialise_parameters(to_dev(normal), seeded_state)
    fast = WindowRunner(CFG, 26, dev).materialise_parameters(to_dev(normal), seeded_state)
    assert fast.operands == "f16"
    # a normal window: no re-run, fp16 and tf32 operands agree to fp32 accuracy
    a, b = fast.infer_window_device(to_dev(normal)).cpu().numpy(), safe.infer_window_device(to_dev(normal)).cpu().numpy()
    assert fast.range_fallbacks == 0
    assert float(np.abs(a - b).max()) <= 1e-4 * float(np.abs(b).max())
    # the out-of-range window: flagged, re-run, same answer as tf32 from the start (same kernels -> tight tolerance)
    want = safe.infer_window_device(to_dev(huge)).cpu().numpy()
    got = fast.infer_window_device(to_dev(huge)).cpu().numpy()
    assert fast.range_fallbacks == 1 and int(fast.range_flag.item()) == 0
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    assert float(np.abs(got[fin] - want[fin]).max()) <= 1e-4 * max(1.0, float(np.abs(want[fin]).max()))
    # graph path: counts_ok() reports the raised flag, infer_window() re-runs eagerly
    graph = GraphWindowRunner(CFG, 26, dev).materialise_parameters(to_dev(normal), seeded_state)
    graph.plan(to_dev(normal))
    graph.capture(to_dev(normal))
    graph.infer_window_device(to_dev(huge))
    assert not graph.counts_ok() and graph.range_fallbacks == 1
    host = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in huge]
    labels = graph.infer_window(host).clone()
    assert (labels.numpy() == want.argmax(1)).mean() > 0.999
    # and a normal window afterwards runs on the graphs again without a re-run
    before = graph.fallbacks
    graph.infer_window([(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in normal])
    assert graph.fallbacks == before
    # several windows in flight: only the lane with the out-of-range window re-runs
    multi = MultiWindowRunner(CFG, 26, dev, lanes=2).prepare(to_dev(normal), seeded_state)
    hn = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in normal]
    out = multi.collect(multi.submit([hn, host]))
    assert (out[1].numpy() == want.argmax(1)).mean() > 0.999
    assert (out[0].numpy() == b.argmax(1)).mean() > 0.999
    assert [l.fallbacks for l in multi.lanes] == [0, 1]


def test_non_default_pointnet_widths_leave_the_graph_path(tmp_path):
    """the static-capacity graphs rely on the FUSED PointNet front end (device-side row counts); a cfg with other widths
    must run eagerly -- and correctly -- instead of reducing uninitialised padding rows (round-1 advisor finding)"""
    import hjson
    import os
    from temporal_latticenet_b200.engine import GraphWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["pointnet_layers"] = [16, 32, 32]
    path = os.path.join(str(tmp_path), "cfg_widths.cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    dev = torch.device("cuda:0")
    w = _window(14, frames=3, n=3000)
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]
    g = GraphWindowRunner(path, 26, dev).materialise_parameters(fd, seeded_state)
    assert not g.supported
    g.capture(fd)
    eager = WindowRunner(path, 26, dev).materialise_parameters(fd, seeded_state)
    want = eager.infer_window_device(fd)
    for rep in range(2):
        got = g.infer_window_device(fd)
        assert len(g.graphs) == 0
        assert float((got - want).abs().max()) < 1e-4 * float(want.abs().max())


def test_submit_captures_a_frame_kind_it_has_not_seen():
    """a 1-frame window (kind first-and-last) submitted to lanes prepared on 4-frame windows (round-1 advisor finding)"""
    from temporal_latticenet_b200.engine import MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    w4 = _window(23, n=4000)
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(to_dev(w4), seeded_state)
    multi = MultiWindowRunner(CFG, 26, dev, lanes=2).prepare(to_dev(w4), seeded_state)
    w1 = w4[:1]
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in (w1, w4)]
    labels = [l.clone() for l in multi.infer_windows(host)]
    assert (True, True) in multi.lanes[0].graphs
    for l, w in zip(labels, (w1, w4)):
        want = eager.infer_window_device(to_dev(w)).argmax(1).cpu()
        assert (l == want).float().mean() > 0.999
    assert sum(l.fallbacks for l in multi.lanes) == 0


def test_lockstep_runner_matches_eager_runner():
    """four windows advancing together, one graph per frame kind for all of them, every tensor-core layer one batched
    launch: same vertex counts and log-softmax (1e-4 of abs-max; only the order of the statistics' atomics differs) as
    running the windows one after the other eagerly; replayed with the windows rotated over the lanes"""
    from temporal_latticenet_b200.engine import LockstepRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    wins = [_window(31), _window(32, n=5200), _window(33, n=4800), _window(34, n=5600)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    devw = [to_dev(w) for w in wins]
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(devw[0], seeded_state)
    ls = LockstepRunner(CFG, 26, dev, lanes=4).prepare(devw[0], seeded_state, devw)
    assert ls.supported and len(ls.graphs) == 3
    assert all(ls.batched[k] > 20 for k in ls.graphs)          # the layers really went through the batched kernel
    want = [eager.infer_window_device(w).cpu().numpy() for w in devw]
    for rep in range(3):
        order = devw[rep:] + devw[:rep]
        outs = ls.infer_windows_device(order)
        torch.cuda.synchronize()
        assert ls.counts_ok()
        for o, w in zip(outs, (want[rep:] + want[:rep])):
            got = o.cpu().numpy()
            assert got.shape == w.shape
            assert float(np.abs(got - w).max()) / float(np.abs(w).max()) < 1e-4
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in wins]
    t1 = ls.submit(host)
    t2 = ls.submit(host[1:] + host[:1])
    for labels, ws in ((ls.collect(t1), want), (ls.collect(t2), want[1:] + want[:1])):
        for l, w in zip(labels, ws):
            assert (l.numpy() == w.argmax(1)).mean() > 0.999
    assert sum(l.fallbacks for l in ls.lanes) == 0
    # a group that does not fill the lanes goes through the lanes' own per-window graphs
    outs = ls.infer_windows_device(devw[:2])
    torch.cuda.synchronize()
    for o, w in zip(outs, want[:2]):
        assert float(np.abs(o.cpu().numpy() - w).max()) / float(np.abs(w).max()) < 1e-4


def test_grouped_lockstep_and_launch_trace():
    """two lock-step groups of two windows in flight together give the eager results; the per-launch trace records of a
    group cover every batched launch with plausible numbers (rows = live vertex counts, durations > 0)"""
    from temporal_latticenet_b200.engine import GroupedLockstepRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    dev = torch.device("cuda:0")
    wins = [_window(41), _window(42, n=5200), _window(43, n=4800), _window(44, n=5600)]
    to_dev = lambda w: [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w]  # noqa: E731
    devw = [to_dev(w) for w in wins]
    eager = WindowRunner(CFG, 26, dev).materialise_parameters(devw[0], seeded_state)
    want = [eager.infer_window_device(w).cpu().numpy() for w in devw]
    g = GroupedLockstepRunner(CFG, 26, dev, lanes=2, groups=2).prepare(devw[0], seeded_state, devw)
    assert g.supported and len(g.lanes) == 4
    for rep in range(2):
        outs = g.infer_windows_device(devw)
        torch.cuda.synchronize()
        assert g.counts_ok()
        for o, w in zip(outs, want):
            assert float(np.abs(o.cpu().numpy() - w).max()) / float(np.abs(w).max()) < 1e-4
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in wins]
    for l, w in zip(g.infer_windows(host), want):
        assert (l.numpy() == w.argmax(1)).mean() > 0.999
    rec = g.trace_group(devw)
    one = g.groups[0]
    assert len(rec) == one.batched_per_group(4) > 80
    assert all(r["us"] > 0 and r["tiles"] > 0 and r["rows"] > 0 and r["ctas"] >= 1 for r in rec)
    v0 = sum(int(l.static_lattice.hash_table.count_tensor().cpu()) for l in one.lanes)
    assert max(r["rows"] for r in rec) == v0          # the last frame's V0-level layers see every vertex of both windows
