"""GPU parity tests of the lattice core: every kernel is driven through the C ABI (via the Python
binding) and compared with the scalar oracle on the same seeded inputs.

Bar: integer structure (keys, vertex ids, neighbour ids) and barycentric weights bit-exact; pure
gathers bit-exact; fp32 accumulations rel 1e-5 / abs 1e-6 (order of summation differs); anything
holding a dense contraction rel 1e-4.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.helpers import REPO, canonical_order, small_window

pytestmark = pytest.mark.gpu

from oracle import lattice_oracle as O  # noqa: E402


def _oracle_funcs():
    path = os.path.join(REPO, "oracle", "shims", "latticenet_py", "lattice", "lattice_funcs.py")
    spec = importlib.util.spec_from_file_location("oracle_lattice_funcs", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _cloud(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        return rng.normal(0, 6, (n, 3)).astype(np.float32), rng.random((n, 1)).astype(np.float32)
    p, v = small_window(seed=seed, frames=1, radius=12.0, max_points=n)[0]
    return p, v


def _cuda_lattice(dev, cap=60000, sigma=0.6):
    from temporal_latticenet_b200.lattice import Lattice
    return Lattice(cap, sigma, device=dev)


@pytest.mark.parametrize("kind,n,sigma", [("gauss", 5000, 0.6), ("lidar", 20000, 0.6), ("gauss", 1, 0.6),
                                           ("gauss", 33, 1.3), ("lidar", 9000, 2.4)])
def test_distribute_matches_oracle(dev, kind, n, sigma):
    pos, val = _cloud(kind, n, 11)
    ls = _cuda_lattice(dev, sigma=sigma)
    rows, idx, w = ls.distribute(torch.from_numpy(pos).to(dev), torch.from_numpy(val).to(dev), True)
    tab = O.OracleTable(60000)
    o_rows, o_idx, o_w = tab.distribute(pos, val, O.scale_factors(sigma))
    V = ls.nr_lattice_vertices()
    assert V == tab.size()
    keys = ls.hash_table.keys().cpu().numpy()
    # canonicalised by sorted key: identical vertex SETS
    assert np.array_equal(keys[canonical_order(keys)], tab.keys()[canonical_order(tab.keys())])
    # and, because numbering is deterministic (order of first appearance), identical ids outright
    assert np.array_equal(keys, tab.keys())
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(w.cpu().numpy(), o_w)  # bit-exact barycentric weights
    o_rows = O.local_mean_sub(o_rows, o_idx, V)
    np.testing.assert_allclose(rows.cpu().numpy(), o_rows, rtol=1e-5, atol=1e-5)
    # size-independent properties
    wsum = w.view(-1, 4).sum(1)
    assert float((wsum - 1).abs().max()) < 1e-5 and float(w.min()) > -1e-6
    cnt = ls.rows_per_vertex(V).cpu().numpy()
    assert cnt.sum() == 4 * pos.shape[0]
    assert np.array_equal(cnt, np.bincount(np.maximum(o_idx, 0), minlength=V).astype(np.float32))


def test_empty_cloud(dev):
    ls = _cuda_lattice(dev)
    rows, idx, w = ls.distribute(torch.zeros(0, 3, device=dev), torch.zeros(0, 1, device=dev), True)
    assert rows.shape == (0, 5) and idx.shape == (0,) and ls.nr_lattice_vertices() == 0


def test_append_only_across_frames_and_reset(dev):
    frames = small_window(seed=5, frames=3, radius=10.0, max_points=8000)
    ls = _cuda_lattice(dev)
    tab = O.OracleTable(60000)
    prev_keys = None
    for f, (pos, val) in enumerate(frames):
        rows, idx, w = ls.distribute(torch.from_numpy(pos).to(dev), torch.from_numpy(val).to(dev), f == 0)
        _, o_idx, _ = tab.distribute(pos, val, O.scale_factors(0.6))
        keys = ls.hash_table.keys().cpu().numpy()
        assert np.array_equal(keys, tab.keys())
        assert np.array_equal(idx.cpu().numpy(), o_idx)
        if prev_keys is not None:  # ids of earlier frames never move (models.py:287-289)
            assert np.array_equal(keys[: prev_keys.shape[0]], prev_keys)
        prev_keys = keys
    # idempotence: inserting a frame again creates nothing
    v = ls.nr_lattice_vertices()
    ls.distribute(torch.from_numpy(frames[-1][0]).to(dev), torch.from_numpy(frames[-1][1]).to(dev), False)
    assert ls.nr_lattice_vertices() == v
    ls.distribute(torch.from_numpy(frames[0][0]).to(dev), torch.from_numpy(frames[0][1]).to(dev), True)
    assert ls.nr_lattice_vertices() < v


def test_capacity_overflow_gives_minus_one(dev):
    """convention U4: vertices beyond the capacity get id -1, in order of first appearance -- exactly
    the oracle's ids while the distinct keys fit the slot array (4 x capacity) ..."""
    pos, val = _cloud("lidar", 9000, 3)
    full = O.OracleTable(60000)
    full.distribute(pos, val, O.scale_factors(0.6))
    cap = int(full.size() * 0.6)
    ls = _cuda_lattice(dev, cap=cap)
    _, idx, _ = ls.distribute(torch.from_numpy(pos).to(dev), torch.from_numpy(val).to(dev), True)
    tab = O.OracleTable(cap)
    _, o_idx, _ = tab.distribute(pos, val, O.scale_factors(0.6))
    assert ls.nr_lattice_vertices() == cap == tab.size()
    assert np.array_equal(idx.cpu().numpy(), o_idx) and (o_idx == -1).any()
    assert ls.hash_table.nr_overflowed() == full.size() - cap
    assert np.array_equal(ls.hash_table.keys().cpu().numpy(), tab.keys())
    # ... and beyond that it degrades without hanging: ids stay in range, the table is exactly full
    pos, val = _cloud("gauss", 4000, 3)
    ls = _cuda_lattice(dev, cap=200)
    _, idx, _ = ls.distribute(torch.from_numpy(pos).to(dev), torch.from_numpy(val).to(dev), True)
    i = idx.cpu().numpy()
    assert ls.nr_lattice_vertices() == 200 and i.max() == 199 and i.min() == -1
    keys = ls.hash_table.keys().cpu().numpy()
    assert len({tuple(k) for k in keys}) == 200


def _three_levels(dev, pos, val):
    ls = _cuda_lattice(dev)
    ls.distribute(torch.from_numpy(pos).to(dev), torch.from_numpy(val).to(dev), True)
    c1 = ls.create_coarse_verts()
    c2 = c1.create_coarse_verts()
    t0, t1, t2 = O.OracleTable(60000), O.OracleTable(60000), O.OracleTable(60000)
    t0.distribute(pos, val, O.scale_factors(0.6))
    t1.insert_points(pos, O.scale_factors(1.2))
    t2.insert_points(pos, O.scale_factors(2.4))
    return (ls, c1, c2), (t0, t1, t2)


def test_coarse_levels_and_neighbour_tables(dev):
    pos, val = _cloud("lidar", 15000, 7)
    (l0, l1, l2), (t0, t1, t2) = _three_levels(dev, pos, val)
    for l, t in ((l0, t0), (l1, t1), (l2, t2)):
        assert np.array_equal(l.hash_table.keys().cpu().numpy(), t.keys())
    for l, t in ((l0, t0), (l1, t1), (l2, t2)):
        for dil in (1, 2):
            assert np.array_equal(l.neighbours(dilation=dil).cpu().numpy(), t.neighbours(dilation=dil))
    assert np.array_equal(l1.neighbours(l0, mode=1).cpu().numpy(), t1.neighbours(t0, mode=1))
    assert np.array_equal(l2.neighbours(l1, mode=1).cpu().numpy(), t2.neighbours(t1, mode=1))
    assert np.array_equal(l0.neighbours(l1, mode=2).cpu().numpy(), t0.neighbours(t1, mode=2))
    assert np.array_equal(l1.neighbours(l2, mode=2).cpu().numpy(), t1.neighbours(t2, mode=2))
    n = l0.neighbours().cpu().numpy()
    V = n.shape[0]
    assert np.array_equal(n[:, 8], np.arange(V))  # slot 8 = centre (lattice_modules.py:320)
    for s in range(8):  # symmetry used by the atomics-free transpose
        m = n[:, s] >= 0
        assert np.array_equal(n[n[m, s], s ^ 1], np.nonzero(m)[0])


@pytest.mark.parametrize("C", [4, 64, 192])
def test_im2row_and_transpose(dev, C):
    from temporal_latticenet_b200 import funcs
    pos, val = _cloud("lidar", 12000, 9)
    (l0, l1, _), (t0, t1, _) = _three_levels(dev, pos, val)
    g = torch.Generator().manual_seed(C)
    for lq, lt, tq, tt, mode in ((l0, l0, t0, t0, 0), (l1, l0, t1, t0, 1), (l0, l1, t0, t1, 2)):
        nbr = tq.neighbours(tt, mode=mode)
        feat = torch.randn(tt.size(), C, generator=g)
        out = funcs.im2row_raw(feat.to(dev), lq.neighbours(lt, mode=mode))
        assert np.array_equal(out.cpu().numpy(), O.im2row(nbr, feat.numpy()))  # pure gather: bit-exact
        # transpose == autograd of the oracle's torch im2row
        of = _oracle_funcs()
        f2 = feat.clone().requires_grad_(True)
        rows = of.im2row_from_table(f2, torch.from_numpy(nbr.astype(np.int64)))
        gr = torch.randn(rows.shape, generator=g)
        rows.backward(gr)
        nbr_t = lt.neighbours(lq, mode={0: 0, 1: 2, 2: 1}[mode])
        got = funcs.row2im_raw(gr.to(dev), nbr_t, tt.size(), C)
        np.testing.assert_allclose(got.cpu().numpy(), f2.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_im2row_short_value_tensor(dev):
    """values may have fewer rows than the lattice has vertices (quirk Q8): missing rows read as 0"""
    from temporal_latticenet_b200 import funcs
    pos, val = _cloud("gauss", 3000, 2)
    (l0, _, _), (t0, _, _) = _three_levels(dev, pos, val)
    feat = torch.randn(t0.size() - 7, 8)
    out = funcs.im2row_raw(feat.to(dev), l0.neighbours())
    assert np.array_equal(out.cpu().numpy(), O.im2row(t0.neighbours(), feat.numpy()))


def test_splat_slice_gather(dev):
    from temporal_latticenet_b200 import funcs
    pos, val = _cloud("lidar", 10000, 4)
    ls = _cuda_lattice(dev)
    p = torch.from_numpy(pos).to(dev)
    feats = torch.randn(pos.shape[0], 3)
    lv, idx, w = funcs.SplatLattice.apply(ls, p, feats.to(dev))
    tab = O.OracleTable(60000)
    _, o_idx, o_w = tab.distribute(pos, np.zeros((pos.shape[0], 1), np.float32), O.scale_factors(0.6))
    V = tab.size()
    np.testing.assert_allclose(lv.cpu().numpy(), O.splat(feats.numpy(), o_idx, o_w, V), rtol=1e-4, atol=1e-4)
    # homogeneous column = sum of barycentric weights per vertex; total mass = number of points
    assert abs(float(lv[:, -1].sum()) - pos.shape[0]) < 1e-2 * pos.shape[0] ** 0.5
    vals = torch.randn(V, 32)
    sl = funcs.SliceLattice.apply(vals.to(dev), ls, p, idx, w)
    np.testing.assert_allclose(sl.cpu().numpy(), O.slice_(vals.numpy(), o_idx, o_w), rtol=1e-5, atol=1e-6)
    # slicing a constant field returns the constant (weights sum to one)
    ones = funcs.SliceLattice.apply(torch.ones(V, 4, device=dev), ls, p, idx, w)
    assert float((ones - 1).abs().max()) < 1e-5
    ga = funcs.GatherLattice.apply(vals[:, :8].contiguous().to(dev), ls, p, idx, w)
    assert np.array_equal(ga.cpu().numpy(), O.gather(vals[:, :8].numpy(), o_idx, o_w))


def test_slice_classify_forward_backward(dev):
    from temporal_latticenet_b200 import funcs
    of = _oracle_funcs()
    pos, val = _cloud("lidar", 6000, 8)
    ls = _cuda_lattice(dev)
    p = torch.from_numpy(pos).to(dev)
    _, idx, w = ls.distribute(p, torch.from_numpy(val).to(dev), True)
    V, C, K, N = ls.nr_lattice_vertices(), 192, 26, pos.shape[0]
    g = torch.Generator().manual_seed(0)
    lv = torch.randn(V, C, generator=g)
    dw = 0.1 * torch.randn(N, 4, generator=g)
    W = torch.randn(K, C, generator=g) / C ** 0.5
    b = torch.randn(K, generator=g)
    ref = O.slice_classify(lv.numpy(), idx.cpu().numpy(), w.cpu().numpy(), dw.numpy(), W.numpy(), b.numpy())
    ins = [t.to(dev).requires_grad_(True) for t in (lv, dw, W, b)]
    out = funcs.SliceClassifyLattice.apply(ins[0], ls, p, ins[1], ins[2], ins[3], K, idx, w)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
    go = torch.randn(N, K, generator=g)
    out.backward(go.to(dev))
    cins = [t.clone().requires_grad_(True) for t in (lv, dw, W, b)]
    cout = of.SliceClassifyLattice.apply(cins[0], None, torch.from_numpy(pos), cins[1], cins[2], cins[3], K, idx.cpu(), w.cpu())
    cout.backward(go)
    for a, c, name in zip(ins, cins, ("lv", "dw", "W", "b")):
        scale = float(c.grad.abs().max()) + 1e-6
        err = float((a.grad.cpu() - c.grad).abs().max()) / scale
        assert err < 1e-4, (name, err)
    # gather / slice backward against autograd of the oracle too
    for fn_name in ("GatherLattice", "SliceLattice"):
        a = lv[:, :16].contiguous().to(dev).requires_grad_(True)
        c = lv[:, :16].clone().requires_grad_(True)
        oa = getattr(funcs, fn_name).apply(a, ls, p, idx, w)
        oc = getattr(of, fn_name).apply(c, None, torch.from_numpy(pos), idx.cpu(), w.cpu())
        gg = torch.randn(oc.shape, generator=g)
        oa.backward(gg.to(dev))
        oc.backward(gg)
        np.testing.assert_allclose(a.grad.cpu().numpy(), c.grad.numpy(), rtol=1e-4, atol=1e-4)
