"""CPU suite (no GPU): pins the oracle, checks the host logic and the C-ABI surface.

 * oracle/window_oracle.py (our restatement of the reference's model recipe) against the goldens the
   reference's OWN models.py / lattice_modules.py produced (tests/golden/make_golden.py);
 * oracle/lattice_oracle.c against the size-independent properties of the permutohedral lattice
    -- the reference ships no vectors for this boundary (parity unpinned there);
 * torch_scatter 2.0.4 semantics of the oracle shim;
 * include/latticenet_b200.h <-> libltn_b200.so symbol agreement (no compute without a GPU).
"""
import ctypes
import json
import os

import hjson
import numpy as np
import pytest
import torch

from tests.helpers import CFG, GOLDEN, REPO, canonical_order, small_window

from oracle import lattice_oracle as O


def _golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return z, json.load(f)


@pytest.mark.parametrize("name", ["lstm_cga_linear_maxpool", "maxpool_aflow_lstm_cga", "aflow_x4", "linear_none_none_gru",
                                  "single_frame", "gru_gru_aflow_gru"])
def test_window_oracle_reproduces_reference_goldens(name, tmp_path):
    from oracle import window_oracle as WO
    z, meta = _golden(name)
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = meta["rnn_modules"]
    cfg["model"]["sequence_learning"] = meta["sequence_learning"]
    path = os.path.join(str(tmp_path), name + ".cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    frames = [(z["pos%d" % i], z["val%d" % i]) for i in range(meta["frames"])]
    torch.manual_seed(0)
    r = WO.OracleWindowRunner(path, meta["nr_classes"])
    r.materialise_parameters(frames)
    shapes = {k: list(v.shape) for k, v in r.model.state_dict().items()}
    assert shapes == meta["shapes"]  # same parameter names and shapes as the reference's state-dict
    got = []
    r.infer_window(frames, collect=got)
    for i in range(meta["frames"]):
        assert got[i][2] == int(z["nv%d" % i])
        if "out%d" % i in z.files:
            np.testing.assert_allclose(got[i][0].numpy(), z["out%d" % i], rtol=1e-4, atol=1e-5)
    assert np.array_equal(r.lattice.hash_table.keys(), z["keys0"])
    np.testing.assert_allclose(got[-1][1].numpy(), z["logits"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("sigma", [0.6, 1.2, 2.4])
def test_simplex_properties(sigma):
    """Sum w = 1, w >= 0, keys sum to zero with a common remainder, sum_i w_i key_i = elevated point"""
    rng = np.random.default_rng(int(sigma * 10))
    pts = rng.normal(0, 20, (2000, 3)).astype(np.float32)
    scale = O.scale_factors(sigma)
    for p in pts:
        keys, bary = O.simplex(p, scale)
        assert abs(float(bary.sum()) - 1.0) < 1e-5 and float(bary.min()) > -1e-6
        full = np.concatenate([keys, -keys.sum(1, keepdims=True)], 1).astype(np.int64)
        for r in range(4):
            assert len(set(np.mod(full[r], 4))) == 1 and np.mod(full[r, 0], 4) == r
        cf = p.astype(np.float64) * scale.astype(np.float64)
        e = np.zeros(4)
        sm = 0.0
        for i in range(3, 0, -1):
            e[i] = sm - i * cf[i - 1]
            sm += cf[i - 1]
        e[0] = sm
        np.testing.assert_allclose((bary[:, None].astype(np.float64) * full).sum(0), e, atol=2e-4 * max(1.0, np.abs(e).max()))


def test_table_append_only_idempotent_overflow_and_levels():
    frames = small_window(seed=5, frames=3, radius=8.0, max_points=4000)
    tab = O.OracleTable(50000)
    scale = O.scale_factors(0.6)
    sizes, prev = [], None
    for p, v in frames:
        rows, idx, w = tab.distribute(p, v, scale)
        keys = tab.keys()
        if prev is not None:
            assert np.array_equal(keys[: prev.shape[0]], prev)  # ids never move (models.py:287-289)
        prev = keys
        sizes.append(tab.size())
        assert idx.min() >= 0 and idx.max() == tab.size() - 1 or idx.max() < tab.size()
        assert np.array_equal(rows[:, -1], w) and np.array_equal(rows[::4, :3], p)
    assert sizes == sorted(sizes)
    tab.distribute(frames[0][0], frames[0][1], scale)
    assert tab.size() == sizes[-1]  # idempotent
    keys = tab.keys()
    assert len({tuple(k) for k in keys}) == keys.shape[0]  # no duplicate vertices
    # neighbour symmetry and the centre slot
    n = tab.neighbours()
    assert np.array_equal(n[:, 8], np.arange(tab.size()))
    for s in range(8):
        m = n[:, s] >= 0
        assert np.array_equal(n[n[m, s], s ^ 1], np.nonzero(m)[0])
    # coarse level: 2*k_coarse is a fine lattice point, so the coarse centre tap may hit a fine vertex
    coarse = O.OracleTable(50000)
    coarse.insert_points(frames[0][0], O.scale_factors(1.2))
    assert 0 < coarse.size() < tab.size()
    nc = coarse.neighbours(tab, mode=1)
    nf = tab.neighbours(coarse, mode=2)
    for s in range(9):  # coarse->fine and fine->coarse tables are transposes of each other
        st = s ^ 1 if s < 8 else 8
        m = nc[:, s] >= 0
        assert np.array_equal(nf[nc[m, s], st], np.nonzero(m)[0])
    # overflow -> -1 (convention U4)
    small = O.OracleTable(100)
    _, idx, _ = small.distribute(frames[0][0], frames[0][1], scale)
    assert small.size() == 100 and (idx == -1).any() and idx.max() == 99


def test_splat_slice_gather_classify_identities():
    p, v = small_window(seed=2, frames=1, radius=7.0, max_points=3000)[0]
    tab = O.OracleTable(50000)
    _, idx, w = tab.distribute(p, v, O.scale_factors(0.6))
    V = tab.size()
    lv = O.splat(v, idx, w, V)
    assert abs(float(lv[:, 1].sum()) - p.shape[0]) < 1e-2  # homogeneous mass = number of points
    const = np.full((V, 4), 3.5, np.float32)
    np.testing.assert_allclose(O.slice_(const, idx, w), 3.5, rtol=1e-5)  # partition of unity
    rng = np.random.default_rng(0)
    vals = rng.normal(size=(V, 8)).astype(np.float32)
    g = O.gather(vals, idx, w).reshape(-1, 4, 9)
    np.testing.assert_allclose(g[:, :, :8].sum(1), O.slice_(vals, idx, w), rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(g[:, :, 8].reshape(-1), w)
    W = rng.normal(size=(5, 8)).astype(np.float32)
    b = rng.normal(size=5).astype(np.float32)
    sc = O.slice_classify(vals, idx, w, np.zeros((p.shape[0], 4), np.float32), W, b)
    np.testing.assert_allclose(sc, O.slice_(vals, idx, w) @ W.T + b, rtol=1e-4, atol=1e-5)
    # linearity in the delta weights
    dw = rng.normal(size=(p.shape[0], 4)).astype(np.float32) * 0.1
    sc2 = O.slice_classify(vals, idx, w, dw, W, b)
    extra = O.slice_(vals, idx, dw.reshape(-1)) @ W.T
    np.testing.assert_allclose(sc2, sc + extra, rtol=1e-3, atol=1e-4)


def test_torch_scatter_oracle_semantics():
    import importlib.util
    spec = importlib.util.spec_from_file_location("orc_scatter", os.path.join(REPO, "oracle", "shims", "torch_scatter", "__init__.py"))
    ts = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ts)
    src = torch.tensor([[1.0, 5.0], [3.0, 5.0], [2.0, -1.0], [7.0, 0.0]])
    idx = torch.tensor([0, 0, 3, 3])
    out, arg = ts.scatter_max(src, idx, dim=0)
    assert out.shape == (4, 2)  # max(index)+1 rows (quirk Q8)
    assert out.tolist() == [[3.0, 5.0], [0.0, 0.0], [0.0, 0.0], [7.0, 0.0]]
    assert arg.tolist() == [[1, 0], [4, 4], [4, 4], [3, 3]]  # tie -> smallest row; empty -> R (Q3 sentinel)
    assert ts.scatter_add(torch.ones(4), idx).tolist() == [2.0, 0.0, 0.0, 2.0]
    assert ts.scatter_mean(src, idx, dim=0)[0].tolist() == [2.0, 5.0]


def test_c_abi_exports_every_declared_symbol():
    from temporal_latticenet_b200 import _lib
    decl = _lib.declared_functions()
    names = [n for n, _ in decl]
    assert len(names) == len(set(names)) and len(names) >= 20
    assert os.path.exists(_lib.LIB_PATH), "build() must have produced the sm_100a library"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    assert lib.ltn_version() >= 100
    # and nothing is exported that the header does not declare
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("ltn_")}
    assert exported == set(names)


def test_product_refuses_to_run_without_cuda():
    """no CPU fallback: the product raises instead of computing anything on the host"""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from temporal_latticenet_b200.lattice import Lattice
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Lattice(1000, 0.6)
    import inspect
    import temporal_latticenet_b200 as pkg
    root = os.path.dirname(inspect.getfile(pkg))
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith(".py"):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "import oracle" not in text and "from oracle" not in text, fn  # the product never touches oracle/


def test_host_logic_cfg_seeding_synthetic():
    from temporal_latticenet_b200.config import ConfigParser
    from temporal_latticenet_b200.lattice import ModelParams, scale_factors
    from temporal_latticenet_b200.seeding import seeded_state
    from temporal_latticenet_b200 import synthetic
    mp = ModelParams.create(CFG)
    assert mp.pointnet_layers() == [16, 32, 64] and mp.nr_downsamples() == 2 and mp.experiment() == "none"
    cp = ConfigParser(CFG)
    assert cp.get_model_vars()["rnn_modules"] == ["gru", "gru", "aflow", "gru"]
    assert cp.get_loader_vars()["frames_per_seq"] == 4
    assert np.array_equal(np.array(scale_factors(0.6), np.float32), O.scale_factors(0.6))  # same bits both sides
    a, b = seeded_state({"x.weight": (3, 4), "y.gn.weight": (5,)}), seeded_state({"y.gn.weight": (5,), "x.weight": (3, 4)})
    assert all(torch.equal(a[k], b[k]) for k in a)
    w = synthetic.World(3)
    p1, v1 = synthetic.scan(w, seed=4, nr_points=5000)
    p2, v2 = synthetic.scan(w, seed=4, nr_points=5000)
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2) and p1.shape == (5000, 3) and p1.dtype == np.float32


def test_canonical_order_helper():
    k = np.array([[2, 0, 0], [0, 1, 0], [0, 0, 5]])
    assert canonical_order(k).tolist() == [2, 1, 0]


def test_scores_oracle_matches_reference_golden():
    """oracle/scores_oracle.py against the outputs of the reference's own callbacks/scores.py (tests/golden/make_scores_golden.py)"""
    import os
    from oracle.scores_oracle import ScoresOracle
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "scores.npz"))
    for case in range(3):
        nr_clouds, K, unl = [int(x) for x in z["c%d_meta" % case]]
        s = ScoresOracle()
        for i in range(nr_clouds):
            s.accumulate_scores(z["c%d_logits%d" % (case, i)], z["c%d_gt%d" % (case, i)], unl)
        assert s.inter == z["c%d_inter" % case].tolist()
        assert s.union == z["c%d_union" % case].tolist()
        avg, per = s.compute_stats()
        assert avg == float(z["c%d_avg" % case])
        assert [per.get(i, -1.0) for i in range(K)] == z["c%d_per" % case].tolist()


def test_lovasz_matches_independent_oracle():
    """the product's vectorised Lovasz-softmax (temporal_latticenet_b200/lovasz.py, plain torch) against the oracle's
    per-class restatement (oracle/shims/latticenet_py/lattice/lovasz_loss.py): value and gradient"""
    import torch
    from oracle import window_oracle  # noqa: F401  (puts oracle/shims first on sys.path)
    from latticenet_py.lattice.lovasz_loss import LovaszSoftmax as OracleLovasz
    from temporal_latticenet_b200.lovasz import LovaszSoftmax
    g = torch.Generator().manual_seed(0)
    for n, k, ignore in ((500, 26, 0), (64, 5, None), (300, 20, 0)):
        logits = torch.randn(n, k, generator=g, dtype=torch.float64)
        target = torch.randint(0, k - 1, (n,), generator=g)   # the last class never occurs: absent classes are skipped
        a = logits.clone().requires_grad_(True)
        b = logits.clone().requires_grad_(True)
        la = LovaszSoftmax(ignore)(torch.log_softmax(a, 1), target)
        lb = OracleLovasz(ignore)(torch.log_softmax(b, 1), target)
        assert abs(float(la) - float(lb)) < 1e-12
        la.backward()
        lb.backward()
        assert float((a.grad - b.grad).abs().max()) < 1e-12
