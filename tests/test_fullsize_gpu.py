"""Parity at BASELINE.json's FULL size (one synthetic SemanticKITTI-shaped window: 4 scans of ~125k points).

What a scalar oracle can still do at this size it does bit-exactly (lattice structure of all three levels,
barycentric weights, neighbour tables: the C oracle needs well under a second); for the feature path, where
the CPU oracle would take minutes per layer stack, size-independent properties stand in: partition of unity
of splat/slice, probabilities summing to one, append-only vertex ids, and the three execution modes (eager,
CUDA-graph replay, several windows in flight) agreeing with each other."""
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
