"""GPU parity of the fused tcgen05 gather-GEMM (csrc/ltn_conv.cu) against a float64 CPU evaluation of
the same contraction over the oracle's neighbour table (oracle im2row).

Tolerance (stated, fp32-parity mode = 3-pass tf32 split): |err| <= 2e-5 * sum_k |a_k||w_k| per output
-- fp32-class (plain fp32 accumulation of K = 576..2304 terms sits at the same level); the single-pass
TF32 variant is checked separately at 2e-3.  The fp16-operand form of the same split (11 + 11 significant
bits per operand in half the bytes, power-of-two scaled, range-flagged) is held to the SAME 2e-5 bound.
"""
import numpy as np
import pytest
import torch

from tests.helpers import small_window

pytestmark = pytest.mark.gpu

from oracle import lattice_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lattice(dev):
    from temporal_latticenet_b200.lattice import Lattice
    p, v = small_window(seed=3, frames=1, radius=14.0, max_points=25000)[0]
    ls = Lattice(60000, 0.6, device=dev)
    ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
    tab = O.OracleTable(60000)
    tab.distribute(p, v, O.scale_factors(0.6))
    assert ls.nr_lattice_vertices() == tab.size()
    return ls, tab


def _check(out, a64, w64, extra64, tol):
    want = a64 @ w64 + extra64
    bound = np.abs(a64) @ np.abs(w64) + np.abs(extra64) + 1e-30
    err = np.abs(out.astype(np.float64) - want) / bound
    assert float(err.max()) < tol, float(err.max())


@pytest.mark.parametrize("C,F", [(64, 64), (128, 128), (192, 192), (256, 128), (64, 256), (32, 16)])
def test_gather_conv_matches_float64(dev, lattice, C, F):
    from temporal_latticenet_b200 import ops
    ls, tab = lattice
    V = tab.size()
    g = torch.Generator().manual_seed(C * 1000 + F)
    x = torch.randn(V, C, generator=g)
    W = torch.randn(9 * C, F, generator=g) / (9 * C) ** 0.5
    nbr = ls.neighbours()
    rows = O.im2row(tab.neighbours(), x.numpy()).astype(np.float64)
    wt = ops.k_major(W.to(dev))
    out = ops.conv_tc(x.to(dev), nbr, wt).cpu().numpy()
    _check(out, rows, W.double().numpy(), np.zeros((V, F)), 2e-5)
    # single-pass TF32 variant, stated separately
    out1 = ops.conv_tc(x.to(dev), nbr, wt, passes=1).cpu().numpy()
    _check(out1, rows, W.double().numpy(), np.zeros((V, F)), 2e-3)
    if C % 64 == 0:   # fp16 hi/lo operands: same bound, and the range flag stays down
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        out16 = ops.conv_tc(x.to(dev), nbr, wt, operands="f16", flag=flag).cpu().numpy()
        _check(out16, rows, W.double().numpy(), np.zeros((V, F)), 2e-5)
        assert int(flag.item()) == 0


def test_folded_groupnorm_relu_bias_residual_and_short_values(dev, lattice):
    from temporal_latticenet_b200 import ops
    ls, tab = lattice
    V, C, F = tab.size(), 128, 64
    g = torch.Generator().manual_seed(7)
    x = torch.randn(V - 37, C, generator=g)          # fewer value rows than vertices (quirk Q8): missing rows read 0
    W = torch.randn(9 * C, F, generator=g) / (9 * C) ** 0.5
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    bias, res = torch.randn(F, generator=g), torch.randn(V, F, generator=g)
    act = torch.relu(x * sc + sh)
    rows = O.im2row(tab.neighbours(), act.numpy()).astype(np.float64)   # absent / missing rows stay exactly 0
    out = ops.conv_tc(x.to(dev), ls.neighbours(), ops.k_major(W.to(dev)), a_scale=sc.to(dev), a_shift=sh.to(dev), relu=True,
                      bias=bias.to(dev), res=res.to(dev)).cpu().numpy()
    _check(out, rows, W.double().numpy(), (bias + res).double().numpy(), 3e-5)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out = ops.conv_tc(x.to(dev), ls.neighbours(), ops.k_major(W.to(dev)), a_scale=sc.to(dev), a_shift=sh.to(dev), relu=True,
                      bias=bias.to(dev), res=res.to(dev), operands="f16", flag=flag).cpu().numpy()
    _check(out, rows, W.double().numpy(), (bias + res).double().numpy(), 3e-5)
    assert int(flag.item()) == 0


@pytest.mark.parametrize("V,C,F", [(1, 64, 64), (127, 64, 48), (129, 96, 16), (5000, 128, 384), (3001, 192, 576), (777, 256, 256)])
def test_dense_rows_and_n_tiling(dev, V, C, F):
    """S = 1 (nn.Linear / GRU gate GEMMs): ragged tile tails, F > 256 split into N tiles"""
    from temporal_latticenet_b200 import ops
    g = torch.Generator().manual_seed(V + C + F)
    x = torch.randn(V, C, generator=g)
    W = torch.randn(F, C, generator=g) / C ** 0.5       # nn.Linear layout = K-major already
    b = torch.randn(F, generator=g)
    out = ops.conv_tc(x.to(dev), None, ops.k_major(W.to(dev), transposed=True), bias=b.to(dev)).cpu().numpy()
    _check(out, x.double().numpy(), W.double().numpy().T, np.broadcast_to(b.double().numpy(), (V, F)), 2e-5)
    if C % 64 == 0:
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        out = ops.conv_tc(x.to(dev), None, ops.k_major(W.to(dev), transposed=True), bias=b.to(dev), operands="f16", flag=flag).cpu().numpy()
        _check(out, x.double().numpy(), W.double().numpy().T, np.broadcast_to(b.double().numpy(), (V, F)), 2e-5)
        assert int(flag.item()) == 0


def test_cross_level_tables(dev, lattice):
    """coarsen (query coarse, values fine) and finefy (query fine, values coarse) through the same kernel"""
    from temporal_latticenet_b200 import ops
    ls, tab = lattice
    coarse = ls.create_coarse_verts()
    tc = O.OracleTable(60000)
    p = ls.positions().cpu().numpy()
    tc.insert_points(p, O.scale_factors(1.2))
    assert coarse.nr_lattice_vertices() == tc.size()
    g = torch.Generator().manual_seed(11)
    C, F = 64, 128
    xf = torch.randn(tab.size(), C, generator=g)
    W = torch.randn(9 * C, F, generator=g) / (9 * C) ** 0.5
    out = ops.conv_tc(xf.to(dev), coarse.neighbours(ls, mode=1), ops.k_major(W.to(dev))).cpu().numpy()
    rows = O.im2row(tc.neighbours(tab, mode=1), xf.numpy()).astype(np.float64)
    _check(out, rows, W.double().numpy(), np.zeros((tc.size(), F)), 2e-5)
    xc = torch.randn(tc.size(), F, generator=g)
    W2 = torch.randn(9 * F, C, generator=g) / (9 * F) ** 0.5
    out = ops.conv_tc(xc.to(dev), ls.neighbours(coarse, mode=2), ops.k_major(W2.to(dev))).cpu().numpy()
    rows = O.im2row(tab.neighbours(tc, mode=2), xc.numpy()).astype(np.float64)
    _check(out, rows, W2.double().numpy(), np.zeros((tab.size(), C)), 2e-5)


@pytest.mark.parametrize("C,F", [(64, 64), (192, 192), (128, 256)])
def test_groupnorm_folded_from_sums_and_output_statistics(dev, lattice, C, F):
    """GN(x)+ReLU folded into the gather from [G,2] sums, and the epilogue's statistics of the output
    (what the next layer's GroupNorm needs) against float64."""
    from temporal_latticenet_b200 import ops
    ls, tab = lattice
    V = tab.size()
    g = torch.Generator().manual_seed(C + F)
    x = torch.randn(V, C, generator=g) * 2 + 0.5
    W = torch.randn(9 * C, F, generator=g) / (9 * C) ** 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    G, Gout = ops.gn_groups(C), ops.gn_groups(F)
    xn = torch.relu(torch.nn.functional.group_norm(x.double().t().unsqueeze(0), G, gamma.double(), beta.double(), 1e-5).squeeze(0).t())
    rows = O.im2row(tab.neighbours(), xn.float().numpy()).astype(np.float64)
    xd = x.to(dev)
    sums = ops.gn_sums(xd, G)
    out_sums = torch.zeros(Gout, 2, dtype=torch.float64, device=dev)
    out = ops.conv_tc(xd, ls.neighbours(), ops.k_major(W.to(dev)), gn=(sums, gamma.to(dev), beta.to(dev), 1e-5), relu=True,
                      out_sums=out_sums).cpu().numpy()
    _check(out, rows, W.double().numpy(), np.zeros((V, F)), 1e-4)   # + the normalisation's own fp32 rounding
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out_sums16 = torch.zeros(Gout, 2, dtype=torch.float64, device=dev)
    out16 = ops.conv_tc(xd, ls.neighbours(), ops.k_major(W.to(dev)), gn=(sums, gamma.to(dev), beta.to(dev), 1e-5), relu=True,
                        out_sums=out_sums16, operands="f16", flag=flag).cpu().numpy()
    _check(out16, rows, W.double().numpy(), np.zeros((V, F)), 1e-4)
    np.testing.assert_allclose(out_sums16.cpu().numpy()[:, 1], out_sums.cpu().numpy()[:, 1], rtol=1e-5)
    assert int(flag.item()) == 0
    cpg = F // Gout
    o64 = out.astype(np.float64).reshape(V, Gout, cpg)
    want = np.stack([o64.sum((0, 2)), (o64 ** 2).sum((0, 2))], 1)
    got = out_sums.cpu().numpy()
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=1e-5)
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=1e-4, atol=1e-3 * np.sqrt(want[:, 1]).max())
    # gn_stats kernel itself
    s64 = x.double().reshape(V, G, C // G)
    np.testing.assert_allclose(sums.cpu().numpy(), np.stack([s64.sum((0, 2)).numpy(), (s64 ** 2).sum((0, 2)).numpy()], 1), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("scale", [1e-3, 1.0, 300.0])
def test_fp16_operands_across_magnitudes_and_range_flag(dev, scale):
    """fp16 hi/lo operands keep the fp32-class bound for small and large activations (power-of-two scaling, subnormal
    lo parts), weights of any magnitude (per-tensor scale), and RAISE the flag -- instead of returning garbage
    silently -- once an activation leaves the representable range (|x| * 2^5 >= 65504)."""
    from temporal_latticenet_b200 import ops
    g = torch.Generator().manual_seed(int(scale * 1000) % 9973)
    V, C, F = 3000, 128, 64
    x = torch.randn(V, C, generator=g) * scale
    for wmag in (1e-4, 1.0, 5e3):
        W = torch.randn(F, C, generator=g) * wmag
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        out = ops.conv_tc(x.to(dev), None, ops.k_major(W.to(dev), transposed=True), operands="f16", flag=flag).cpu().numpy()
        _check(out, x.double().numpy(), W.double().numpy().T, np.zeros((V, F)), 2e-5)
        assert int(flag.item()) == 0
    xb = x.clone()
    xb[1234, 77] = 2100.0      # 2100 * 32 > 65504
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.conv_tc(xb.to(dev), None, ops.k_major(W.to(dev), transposed=True), operands="f16", flag=flag)
    assert int(flag.item()) == 1
    # the tf32 operands take the same input without a flag
    out = ops.conv_tc(xb.to(dev), None, ops.k_major(W.to(dev), transposed=True)).cpu().numpy()
    _check(out, xb.double().numpy(), W.double().numpy().T, np.zeros((V, F)), 2e-5)


def test_alternative_operand_paths_agree(dev, lattice):
    """The two experimental operand paths (A staged in tensor memory for a TS-mode MMA; weight tiles multicast across
    a thread-block cluster) are kept behind environment switches because they measured slower; they must still be
    correct.  Each switch is read once per process, so they run in subprocesses."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from temporal_latticenet_b200 import ops\n"
        "from temporal_latticenet_b200.lattice import Lattice\n"
        "from tests.helpers import small_window\n"
        "dev = torch.device('cuda:0')\n"
        "p, v = small_window(seed=3, frames=1, radius=14.0, max_points=25000)[0]\n"
        "ls = Lattice(60000, 0.6, device=dev)\n"
        "ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)\n"
        "V = ls.nr_lattice_vertices(); g = torch.Generator().manual_seed(5)\n"
        "for C, F in ((64, 64), (192, 192), (128, 48)):\n"
        "    x = torch.randn(V, C, generator=g); W = torch.randn(9 * C, F, generator=g) / (9 * C) ** 0.5\n"
        "    nbr = ls.neighbours().cpu().long(); pad = torch.cat([x, torch.zeros(1, C)], 0).double()\n"
        "    rows = pad[torch.where(nbr < 0, torch.full_like(nbr, V), nbr).reshape(-1)].reshape(V, 9 * C)\n"
        "    want = rows @ W.double(); bound = rows.abs() @ W.double().abs() + 1e-30\n"
        "    flag = torch.zeros(1, dtype=torch.int32, device=dev)\n"
        "    for mode in ('tf32', 'f16'):\n"
        "        out = ops.conv_tc(x.to(dev), ls.neighbours(), ops.k_major(W.to(dev)), operands=mode, flag=flag).cpu().double()\n"
        "        assert float(((out - want).abs() / bound).max()) < 2e-5, (C, F, mode)\n"
        "    assert int(flag.item()) == 0\n"
        "print('ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # tf32: A in tensor memory (register transpose); fp16: A through the shared-memory ring instead of tensor memory;
    # weight tiles multicast across clusters of 4 / 2 (both operand types)
    for env in ({"LTN_CONV_ATMEM": "1"}, {"LTN_CONV_ATMEM": "0"}, {"LTN_CONV_CLUSTER": "4"}, {"LTN_CONV_CLUSTER": "2"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, (env, r.stdout[-500:], r.stderr[-1500:])


# ---------------------------------------------------------------------------------------------------------------
# Bit reproducibility: k_conv_tc writes `out` without atomics (one CTA owns each output tile, fixed k order), so the
# same inputs must give the SAME BITS on every launch -- alone and with four streams hammering the kernel at once.
# Every convolution shape of the KITTI model, both operand types, with and without the folded GroupNorm.  (The
# epilogue's GroupNorm statistics ARE accumulated with atomics and are compared with a tolerance instead.)
# ---------------------------------------------------------------------------------------------------------------
MODEL_SHAPES = [  # (C, F, S): SURVEY.md B.10 -- convs, coarsen / finefy, 1x1 layers, GRU / AFlow dense layers
    (128, 64, 9), (64, 64, 9), (64, 128, 9), (128, 128, 9), (128, 256, 9), (256, 128, 9), (192, 192, 9),
    (256, 64, 1), (64, 256, 1), (192, 192, 1), (192, 96, 1), (128, 384, 1), (64, 192, 1), (192, 576, 1), (512, 256, 1)]


@pytest.mark.parametrize("operands", ["f16", "tf32"])
def test_conv_tc_is_bit_reproducible(dev, lattice, operands):
    from temporal_latticenet_b200 import ops
    ls, tab = lattice
    V = tab.size()
    nbr = ls.neighbours()
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    reps = 40
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    for C, F, S in MODEL_SHAPES:
        if operands == "f16" and C % 64:
            continue
        g = torch.Generator().manual_seed(C * 7 + F * 3 + S)
        x = torch.randn(V, C, generator=g).to(dev)
        W = (torch.randn(S * C, F, generator=g) / (S * C) ** 0.5).to(dev)
        wt = ops.k_major(W)
        gamma, beta = torch.rand(C, generator=g).to(dev) + 0.5, torch.randn(C, generator=g).to(dev)
        for folded in ((False, True) if C <= 256 else (False,)):
            gn = (ops.gn_sums(x, ops.gn_groups(C)), gamma, beta, 1e-5) if folded else None
            kw = dict(gn=gn, relu=folded, operands=operands, flag=flag)
            first = ops.conv_tc(x, nbr if S == 9 else None, wt, **kw)
            torch.cuda.synchronize()
            for _ in range(reps):
                again = ops.conv_tc(x, nbr if S == 9 else None, wt, **kw)
                assert torch.equal(first, again), (C, F, S, folded, "sequential")
            # four streams at once, each with its own output (and statistics) buffer
            outs, sums = [], []
            cur = torch.cuda.current_stream()
            for s in streams:
                s.wait_stream(cur)
            for r in range(reps // 4):
                for s in streams:
                    with torch.cuda.stream(s):
                        sm = torch.zeros(ops.gn_groups(F), 2, dtype=torch.float64, device=dev)
                        outs.append(ops.conv_tc(x, nbr if S == 9 else None, wt, out_sums=sm, **kw))
                        sums.append(sm)
            for s in streams:
                cur.wait_stream(s)
            torch.cuda.synchronize()
            want_sums = torch.stack([first.double().view(V, ops.gn_groups(F), -1).sum((0, 2)),
                                     (first.double() ** 2).view(V, ops.gn_groups(F), -1).sum((0, 2))], 1)
            for o, sm in zip(outs, sums):
                assert torch.equal(first, o), (C, F, S, folded, "concurrent")
                assert torch.allclose(sm, want_sums, rtol=1e-5, atol=1e-3), (C, F, S, folded, "statistics")
    assert int(flag.item()) == 0


# ---------------------------------------------------------------------------------------------------------------
# Batched persistent kernel (csrc/ltn_conv_batched.cu): one launch for the same layer of several windows must give
# the SAME BITS as one ltn_conv_tc_f16 launch per window (same MMAs, same k order per tile), for every model shape,
# ragged problem sizes, device-side row counts, folded GroupNorm, bias, residual, output statistics.
# ---------------------------------------------------------------------------------------------------------------
class _Collector:
    def __init__(self):
        self.reqs = []

    def request(self, d):
        self.reqs.append(d)


def _batched_vs_single(dev, ls, V, shape, nb, folded, with_res, with_bias, with_sums, seed, dev_counts=False):
    from temporal_latticenet_b200 import ops
    C, F, S = shape
    nbr_full = ls.neighbours()
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(S * C, F, generator=g) / (S * C) ** 0.5).to(dev)
    wt = ops.k_major(W)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(dev), torch.randn(C, generator=g).to(dev)
    bias = torch.randn(F, generator=g).to(dev) if with_bias else None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    probs = []
    for b in range(nb):
        Vb = max(1, V - 137 * b - (b * V) // 7)                       # ragged: every problem has its own row count
        x = torch.randn(V, C, generator=g).to(dev) * (1.0 + 0.3 * b)
        nbr = nbr_full[:Vb].contiguous() if S == 9 else None
        res = torch.randn(Vb, F, generator=g).to(dev) if with_res else None
        live = Vb - 5 * b if dev_counts else Vb                        # device-side count below the host bound
        cnt = torch.tensor([live], dtype=torch.int32, device=dev) if dev_counts else None
        probs.append((x, nbr, res, Vb, live, cnt))
    import ctypes

    def run(batched):
        outs, sums = [], []
        col = _Collector()
        ops._BATCH.ctx = col if batched else None
        try:
            for x, nbr, res, Vb, live, cnt in probs:
                gn = (ops.gn_sums(x, ops.gn_groups(C)), gamma, beta, 1e-5) if folded else None
                sm = torch.zeros(ops.gn_groups(F), 2, dtype=torch.float64, device=dev) if with_sums else None
                out = torch.zeros(Vb, F, device=dev)
                rd = ctypes.c_void_p(cnt.data_ptr()) if cnt is not None else None
                o = ops.conv_tc(x, nbr, wt, nr_rows=Vb, gn=gn, relu=folded, bias=bias, res=res, out=out, out_sums=sm, operands="f16",
                                flag=flag, rows_dev=rd if S == 1 else None)
                if cnt is not None and S == 9:   # gather form: the count bounds the QUERY rows only
                    pass
                outs.append(o)
                sums.append(sm)
        finally:
            ops._BATCH.ctx = None
        if batched:
            assert len(col.reqs) == nb
            if dev_counts and S == 9:
                for r, (_, _, _, _, _, cnt) in zip(col.reqs, probs):
                    r["vq_dev"] = ctypes.c_void_p(cnt.data_ptr())
            ops.conv_tc_batched(col.reqs)
        torch.cuda.synchronize()
        return outs, sums
    if dev_counts and S == 9:
        return None   # exercised through the S == 1 form (rows_dev) and the engine tests
    a, sa = run(False)
    b, sb = run(True)
    for i, (o1, o2) in enumerate(zip(a, b)):
        live = probs[i][4]
        assert torch.equal(o1[:live], o2[:live]), (shape, nb, i, "out differs")
        if with_sums:
            assert torch.allclose(sa[i], sb[i], rtol=1e-6, atol=1e-3), (shape, nb, i, "statistics differ")
    assert int(flag.item()) == 0


@pytest.mark.parametrize("nb", [1, 3, 4, 8])
def test_batched_conv_is_bit_identical_to_single_launches(dev, lattice, nb):
    ls, tab = lattice
    V = tab.size()
    for k, shape in enumerate(MODEL_SHAPES):
        C, F, S = shape
        if C % 64:
            continue
        folded = C <= 256 and k % 2 == 0
        _batched_vs_single(dev, ls, V, shape, nb, folded, with_res=(k % 3 == 0), with_bias=(k % 2 == 1), with_sums=(F % 32 == 0 and k % 4 != 3),
                           seed=100 * nb + k)


def test_batched_conv_device_side_counts_and_tiny_problems(dev, lattice):
    ls, tab = lattice
    V = tab.size()
    for shape in [(64, 64, 1), (128, 384, 1), (192, 192, 1), (64, 192, 1)]:
        _batched_vs_single(dev, ls, V, shape, 4, folded=True, with_res=True, with_bias=True, with_sums=shape[1] % 32 == 0, seed=7,
                           dev_counts=True)
    # problems smaller than one tile, and a group in which most CTAs have nothing to do
    _batched_vs_single(dev, ls, 150, (64, 64, 9), 3, folded=True, with_res=False, with_bias=False, with_sums=True, seed=9)
    _batched_vs_single(dev, ls, 40, (192, 192, 9), 2, folded=False, with_res=True, with_bias=True, with_sums=True, seed=10)


@pytest.mark.parametrize("kind", ["conv_gn_res_bias", "conv_plain", "linear_gn_res_bias", "coarsen_gn"])
def test_fused_training_layer_gradients_match_unfused(dev, lattice, kind):
    """funcs._FusedConv (training: the fused tensor-core kernel under autograd; backward = act recompute + the forward kernel
    over the transposed table + gathered-act^T . dy + the GroupNorm/ReLU backward kernel) against the UNFUSED autograd
    composition (ops.group_norm -> funcs.gather_conv / torch linear -> + bias -> + res) on the same inputs: output and every
    gradient (x, weight, bias, residual, gamma, beta) within 2e-4 of the abs-max (fp32 summation orders differ)."""
    from temporal_latticenet_b200 import funcs, ops
    from temporal_latticenet_b200.modules import GroupNormLatticeModule
    ls, _ = lattice
    g = torch.Generator().manual_seed(11)
    V = ls.nr_lattice_vertices()
    C, F = (128, 64) if kind != "linear_gn_res_bias" else (64, 256)
    linear = kind.startswith("linear")
    with_gn = "gn" in kind
    with_rb = "res_bias" in kind
    if kind == "coarsen_gn":
        coarse = ls.create_coarse_verts()
        nbr, nbr_t = coarse.neighbours(ls, mode=1), ls.neighbours(coarse, mode=2)
    else:
        nbr = nbr_t = ls.neighbours()
    Vq = V if linear else nbr.shape[0]
    ops.begin_frame(ls, dev)

    def leaf(*shape, scale=1.0, shift=0.0):
        return (torch.randn(*shape, generator=g) * scale + shift).to(dev).requires_grad_(True)
    x0 = leaf(V, C, scale=2.0, shift=0.3)
    w0 = leaf(F, C, scale=C ** -0.5) if linear else leaf(9 * C, F, scale=(9 * C) ** -0.5)
    b0 = leaf(F) if with_rb else None
    r0 = leaf(Vq, F) if with_rb else None
    norm = GroupNormLatticeModule(C) if with_gn else None
    if norm is not None:
        with torch.no_grad():
            norm.gn.weight.copy_(torch.rand(C, generator=g) + 0.5)
            norm.gn.bias.copy_(torch.randn(C, generator=g) * 0.2)
    gy = torch.randn(Vq, F, generator=g).to(dev)

    def grads(out, leaves):
        for t in leaves:
            if t is not None and t.grad is not None:
                t.grad = None
        (out * gy).sum().backward()
        return [None if t is None else t.grad.clone() for t in leaves]
    leaves = [x0, w0, b0, r0] + ([norm.gn.weight, norm.gn.bias] if norm is not None else [])
    fused = funcs.fused_conv_train(x0, w0, b0, r0, norm, None if linear else nbr, None if linear else (lambda: nbr_t), linear=linear)
    gf = grads(fused, leaves)
    a = ops.group_norm(x0, norm.gn.weight, norm.gn.bias, norm.groups, norm.gn.eps, True) if norm is not None else x0
    plain = torch.nn.functional.linear(a, w0) if linear else funcs.gather_conv(a, w0, nbr, lambda: nbr_t)
    if with_rb:
        plain = plain + b0 + r0
    gp = grads(plain, leaves)

    def rel(u, v):
        return float((u - v).abs().max()) / (float(v.abs().max()) + 1e-30)
    assert rel(fused.detach(), plain.detach()) < 2e-4
    for name, u, v in zip(("x", "weight", "bias", "res", "gamma", "beta"), gf, gp):
        if v is not None:
            assert rel(u, v) < 2e-4, (kind, name, rel(u, v))


@pytest.mark.parametrize("C,F,S", [(64, 64, 9), (192, 192, 9), (128, 256, 9), (256, 128, 9), (96, 16, 1), (192, 96, 1), (64, 256, 1), (36, 48, 1)])
def test_weight_gradient_on_tensor_cores_matches_float64(dev, lattice, C, F, S):
    """csrc/ltn_conv_bwd_weight.cu (both operands MN-major in shared memory, 3-pass tf32 split, split-K over vertex ranges with
    fp32 reductions) against the same contraction in float64 over the oracle's neighbour table:
    dW[s*C + c, f] = sum_v act[nbr[v,s], c] * dy[v, f].  Bound as for the forward kernel: 2e-5 * sum|a||dy| per entry."""
    from temporal_latticenet_b200 import funcs
    ls, tab = lattice
    V = ls.nr_lattice_vertices()
    g = torch.Generator().manual_seed(C + F + S)
    act = torch.relu(torch.randn(V, C, generator=g)) * 1.5
    dy = torch.randn(V, F, generator=g) * 1e-3
    nbr = ls.neighbours() if S == 9 else None
    got = funcs.conv_bwd_weight(act.to(dev), dy.to(dev), nbr)
    assert got is not None and tuple(got.shape) == (S * C, F)
    got = got.cpu().numpy().astype(np.float64)
    a64, d64 = act.numpy().astype(np.float64), dy.numpy().astype(np.float64)
    rows = O.im2row(tab.neighbours(), act.numpy()).astype(np.float64) if S == 9 else a64   # [V, S*C]
    want = rows.T @ d64
    bound = np.abs(rows).T @ np.abs(d64) + 1e-30
    assert float((np.abs(got - want) / bound).max()) < 2e-5, float((np.abs(got - want) / bound).max())
