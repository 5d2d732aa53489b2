"""GPU parity of the per-vertex temporal fusion kernels (seq_lattice/lattice_modules.py:17-339) and
the segmented reductions / GroupNorm against plain torch fp32 on the CPU (floating-point kernels:
torch is the checker here, the goldens in test_golden_gpu.py pin the reference's own module code).

Tolerances: pointwise gate kernels rel 1e-5 (expf/tanhf vs the CPU's libm differ in the last ulp);
AFlow rel 1e-4 (9 x C-term distances, sqrt, division); GroupNorm 1e-5 abs on unit-variance data.
"""
import numpy as np
import pytest
import torch

from tests.helpers import small_window

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _lattice_two_frames(dev):
    from temporal_latticenet_b200.lattice import Lattice
    frames = small_window(seed=8, frames=2, radius=9.0, max_points=5000)
    ls = Lattice(60000, 1.2, device=dev)
    counts = []
    for f, (p, v) in enumerate(frames):
        ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), f == 0)
        counts.append(ls.nr_lattice_vertices())
    return ls, counts


@pytest.mark.parametrize("C", [64, 128, 192])
def test_gru_and_lstm_fused_equal_torch_cells(dev, C):
    from temporal_latticenet_b200.fusion import GRUModule, LSTMModule
    ls, (vh, v) = _lattice_two_frames(dev)
    g = torch.Generator().manual_seed(C)
    h0, x1 = torch.randn(vh, C, generator=g), torch.randn(v, C, generator=g)
    for cls, cell_name in ((GRUModule, "GRU"), (LSTMModule, "lstm")):
        torch.manual_seed(1)
        m = cls(C)
        m_cpu_state = {k: t.clone() for k, t in m.state_dict().items()}
        m = m.to(dev)
        with torch.no_grad():
            m(h0.to(dev), ls)
            out, _ = m(x1.to(dev), ls)
        # the reference's recipe, literally, on the CPU (lattice_modules.py:27-37,53-63)
        ref = cls(C)
        ref.load_state_dict(m_cpu_state)
        with torch.no_grad():
            h = ref.hidden_linear(h0)
            h = torch.nn.utils.rnn.pad_sequence([h, x1], padding_value=0.0)[:, 0, :]
            if cell_name == "GRU":
                want = ref.GRU(x1, h)
            else:
                want, _ = ref.lstm(x1, (h, torch.zeros_like(h)))
        np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5)
        # and the differentiable path (used under autograd) gives the same numbers
        m.reset_sequence()
        xg = x1.to(dev).requires_grad_(True)
        m(h0.to(dev), ls)
        out2, _ = m(xg, ls)
        np.testing.assert_allclose(out2.detach().cpu().numpy(), out.cpu().numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("C,use_center", [(64, True), (256, True), (128, False)])
def test_aflow_fused_equals_reference_recipe(dev, C, use_center):
    """lattice_modules.py:298-339 restated with torch on the CPU over the oracle's neighbour table"""
    from temporal_latticenet_b200.fusion import CustomKernelConvLatticeIm2RowModule
    ls, (vh, v) = _lattice_two_frames(dev)
    g = torch.Generator().manual_seed(C)
    h, lv = torch.randn(vh, C, generator=g), torch.randn(v, C, generator=g)
    m = CustomKernelConvLatticeIm2RowModule(C, use_center=use_center).to(dev)
    with torch.no_grad():
        out, w, nbr = m(lv.to(dev), h.to(dev), ls)
    bias = m.bias.detach().cpu()
    nbr_c = nbr.cpu().long()
    hp = torch.nn.functional.pad(h, (0, 0, 0, v - vh), value=-999999.0)
    pad = torch.cat([hp, torch.zeros(1, C)], 0)
    nb = pad[torch.where(nbr_c < 0, torch.full_like(nbr_c, v), nbr_c)]  # [V,9,C]
    present = (nbr_c != -1).float()
    d = torch.cdist(nb, lv.unsqueeze(1), p=2.0).squeeze(2) * present
    if not use_center:
        d[:, 8] = 0
    d = d / d.sum(1, keepdim=True)
    a = torch.ones_like(d) * 0.1
    ww = (a - torch.min(d, a)) * 0.1 * present
    if not use_center:
        ww[:, 8] = 0
    want = (nb * ww.unsqueeze(2)).sum(1) + bias
    got, gw = out.cpu(), w.cpu()
    assert torch.equal(torch.isnan(gw), torch.isnan(ww))  # quirk Q5: NaNs where the reference makes them
    ok = ~torch.isnan(ww).any(1)
    np.testing.assert_allclose(gw[ok].numpy(), ww[ok].numpy(), rtol=1e-4, atol=1e-6)
    scale = float(want[ok].abs().max())
    assert float((got[ok] - want[ok]).abs().max()) < 1e-4 * scale
    # differentiable path == fused path
    lvg = lv.to(dev).requires_grad_(True)
    out2, w2, _ = m(lvg, h.to(dev), ls)
    assert float((out2.detach().cpu()[ok] - got[ok]).abs().max()) < 1e-4 * scale


@pytest.mark.parametrize("V,C", [(1000, 64), (4097, 192), (300, 36), (5, 8)])
def test_group_norm_relu(dev, V, C):
    from temporal_latticenet_b200 import ops
    g = torch.Generator().manual_seed(V)
    x = torch.randn(V, C, generator=g) * 3 + 1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    G = ops.gn_groups(C)
    xg = x.to(dev).requires_grad_(True)
    gg, bg = gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
    y = ops.group_norm(xg, gg, bg, G, 1e-5, True)
    xc, gc, bc = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    yc = torch.relu(torch.nn.functional.group_norm(xc.t().unsqueeze(0), G, gc, bc, 1e-5).squeeze(0).t())
    np.testing.assert_allclose(y.detach().cpu().numpy(), yc.detach().float().numpy(), rtol=1e-5, atol=1e-5)
    go = torch.randn(V, C, generator=g)
    y.backward(go.to(dev))
    yc.backward(go.double())
    for a, b in ((xg, xc), (gg, gc), (bg, bc)):
        s = float(b.grad.abs().max()) + 1e-9
        assert float((a.grad.cpu().double() - b.grad).abs().max()) / s < 1e-4


def test_scatter_max_add_mean_semantics(dev):
    """torch_scatter 2.0.4 semantics at the reference's call sites (lattice_modules.py:485-520):
    empty segment -> 0 / argmax = R; ids < 0 folded by the caller; smallest row wins ties."""
    from temporal_latticenet_b200 import ops
    g = torch.Generator().manual_seed(3)
    R, C, V = 5000, 64, 700
    src = torch.randn(R, C, generator=g)
    src[10] = src[3]  # a tie between rows 3 and 10 when they share a segment
    idx = torch.randint(0, V - 50, (R,), generator=g)  # last 50 segments empty
    idx[10] = idx[3]
    out, arg = ops.scatter_max(src.to(dev), idx.to(dev), dim_size=V)
    want = torch.full((V, C), float("-inf")).scatter_reduce(0, idx.view(-1, 1).expand(R, C), src, "amax")
    empty = torch.isinf(want)
    want = torch.where(empty, torch.zeros_like(want), want)
    assert torch.equal(out.cpu(), want)
    a = arg.cpu()
    assert bool((a[empty] == R).all())
    assert torch.equal(src.gather(0, a.clamp(max=R - 1))[~empty], want[~empty])
    seg = int(idx[3])
    tie_cols = (a[seg] == 10)
    assert not bool(tie_cols.any())  # row 3 beats row 10 on equal values
    add = ops.scatter_add(src.to(dev), idx.to(dev), dim_size=V).cpu()
    want_add = torch.zeros(V, C).index_add(0, idx, src)
    np.testing.assert_allclose(add.numpy(), want_add.numpy(), rtol=1e-4, atol=1e-4)
    mean = ops.scatter_mean(src.to(dev), idx.to(dev), dim_size=V).cpu()
    cnt = torch.bincount(idx, minlength=V).clamp(min=1).float().unsqueeze(1)
    np.testing.assert_allclose(mean.numpy(), (want_add / cnt).numpy(), rtol=1e-4, atol=1e-4)
    # differentiable: gradient lands on the arg-max rows only
    s = src.to(dev).requires_grad_(True)
    o, ar = ops.scatter_max(s, idx.to(dev), dim_size=V)
    o.sum().backward()
    hits = torch.zeros(R, C)
    hits.scatter_add_(0, a.clamp(max=R - 1), (~empty).float())
    assert torch.equal(s.grad.cpu(), hits)


@pytest.mark.parametrize("order", ["scan", "shuffled"])
@pytest.mark.parametrize("kernel", ["cuda_cores", "tensor_cores"])
def test_pointnet_front_end_matches_float64(dev, kernel, order):
    """PointNetSeqModule front end (lattice_modules.py:448-530): MLP 4 -> 16 -> 32 -> 64 per distributed row, segmented
    max per vertex with arg-max (smallest row on ties), barycentric weight of the winning row (quirk Q3), min-4-rows mask.
    Both kernels -- the whole MLP on the CUDA cores, and the 32 -> 64 layer on the tensor cores with fp16 hi/lo operands --
    against a float64 evaluation: maxima within 1e-5 (relative to the layer's |a||w| sum), and where the float64 runner-up
    is not within that tolerance of the maximum the arg-max row, hence the gathered barycentric weight, must be the same."""
    from temporal_latticenet_b200 import _lib
    from temporal_latticenet_b200.lattice import Lattice
    p_np, v_np = small_window(seed=5, frames=1, radius=12.0, max_points=20000)[0]
    if order == "shuffled":   # the reference shuffles the points of a training scan (kitti_dataloader.py:174-180): a block of
        perm = np.random.default_rng(1).permutation(p_np.shape[0])   # rows then touches many distinct vertices, which takes
        p_np, v_np = np.ascontiguousarray(p_np[perm]), np.ascontiguousarray(v_np[perm])   # the kernels' table-overflow path
    ls = Lattice(60000, 0.6, device=dev)
    rows, idx, w = ls.distribute(torch.from_numpy(p_np).to(dev), torch.from_numpy(v_np).to(dev), True)
    V, R = ls.nr_lattice_vertices(), rows.shape[0]
    g = torch.Generator().manual_seed(3)
    w1, b1 = torch.randn(16, 4, generator=g) * 0.7, torch.randn(16, generator=g) * 0.3
    w2, b2 = torch.randn(32, 16, generator=g) * 0.35, torch.randn(32, generator=g) * 0.3
    w3, b3 = torch.randn(64, 32, generator=g) * 0.25, torch.randn(64, generator=g) * 0.3
    P, lib = _lib.ptr, _lib.load()
    d = [t.to(dev).contiguous() for t in (w1, b1, w2, b2, w3, b3)]
    packed = torch.empty(V, 64, dtype=torch.int64, device=dev)
    out = torch.empty(V, 128, device=dev)
    if kernel == "cuda_cores":
        rc = lib.ltn_pointnet(P(rows), 5, P(idx), R, None, *[P(t) for t in d], V, None, P(packed), P(ls._vert_acc), 4, P(out), _lib.stream())
    else:
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        w12 = np.ascontiguousarray(np.concatenate([t.numpy().reshape(-1) for t in (w1, b1, w2, b2)]).astype(np.float32))
        import ctypes
        rc = lib.ltn_pointnet_tc(P(rows), 5, P(idx), R, None, w12.ctypes.data_as(ctypes.c_void_p), P(d[4]), P(d[5]), V, None, P(packed), P(ls._vert_acc), 4, P(out), 5, P(flag),
                                 _lib.stream())
        assert int(flag.item()) == 0
    assert rc == 0
    got = out.cpu().double().numpy()
    # float64 reference
    x = rows.cpu().double()
    ids = idx.cpu().long().clamp(min=0)
    h = torch.relu(x[:, :4] @ w1.double().t() + b1.double())
    h = torch.relu(h @ w2.double().t() + b2.double())
    y = h @ w3.double().t() + b3.double()
    bound = (h.abs() @ w3.double().abs().t() + b3.double().abs()).numpy()
    y, ids = y.numpy(), ids.numpy()
    bary = x[:, 4].numpy()
    cnt = np.bincount(ids, minlength=V)
    order = np.argsort(ids, kind="stable")
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    checked = 0
    for v in range(V):
        if cnt[v] < 4:
            assert not got[v].any()
            continue
        rws = order[starts[v]:starts[v] + cnt[v]]
        yy = y[rws]
        best = yy.max(0)
        tol = 1e-5 * bound[rws].max(0)
        assert np.all(np.abs(got[v, :64] - best) <= tol + 1e-7), (v, np.abs(got[v, :64] - best).max())
        arg = rws[yy.argmax(0)]
        arg = np.where(arg > V, 0, arg)                      # quirk Q3, literally
        srt = np.sort(yy, 0)
        clear = (srt[-1] - srt[-2] > 4 * tol) if cnt[v] > 1 else np.ones(64, bool)
        assert np.array_equal(got[v, 64:][clear], bary[arg][clear].astype(np.float32).astype(np.float64)), v
        checked += int(clear.sum())
    assert checked > 32 * (cnt >= 4).sum()


def test_scores_iou_matches_oracle():
    """device-side IoU accumulation (callbacks/scores.py:13-47) against the pinned numpy oracle, on the golden clouds
    and on one full-size cloud; the result also goes through the reference-named accessors"""
    import os
    from oracle.scores_oracle import ScoresOracle
    from temporal_latticenet_b200.scores import Scores
    from tests.helpers import GOLDEN
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(GOLDEN, "scores.npz"))
    for case in range(3):
        nr_clouds, K, unl = [int(x) for x in z["c%d_meta" % case]]
        s = Scores()
        for i in range(nr_clouds):
            s.accumulate_scores(torch.from_numpy(z["c%d_logits%d" % (case, i)]).to(dev), torch.from_numpy(z["c%d_gt%d" % (case, i)]).to(dev), unl)
        assert s.intersection_per_class == z["c%d_inter" % case].tolist()
        assert s.union_per_class == z["c%d_union" % case].tolist()
        assert s.avg_class_iou() == float(z["c%d_avg" % case])
        per = s.iou_per_class()
        assert [per.get(i, -1.0) for i in range(K)] == z["c%d_per" % case].tolist()
        s.update_best()
        assert s.best_iou == float(z["c%d_avg" % case])
    rng = np.random.default_rng(5)
    n, K = 125000, 26
    logits = rng.standard_normal((n, K)).astype(np.float32)
    gt = rng.integers(0, K, n)
    logits[np.arange(n), gt] += 1.0
    s, o = Scores(), ScoresOracle()
    for rep in range(2):
        s.accumulate_scores(torch.from_numpy(logits).to(dev), torch.from_numpy(gt).to(dev), 0)
        o.accumulate_scores(logits, gt, 0)
    assert s.intersection_per_class == o.inter and s.union_per_class == o.union
    assert s.compute_stats() == o.compute_stats()


@pytest.mark.parametrize("C,relu", [(64, True), (192, True), (36, True), (128, False), (8, False)])
def test_group_norm_backward_kernel_matches_torch_autograd(C, relu):
    """ltn_gn_bwd (two kernels) against torch.nn.functional.group_norm (+ReLU) under autograd in float64; also the conv
    backward-data path through the tensor-core kernel over the transposed table is covered by test_golden_gpu's gradient test"""
    from temporal_latticenet_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C)
    V = 3001
    x = torch.randn(V, C, generator=g).to(dev)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(dev), torch.randn(C, generator=g).to(dev)
    gy = torch.randn(V, C, generator=g).to(dev)
    G = ops.gn_groups(C)
    xa, ga, ba = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ops.group_norm(xa, ga, ba, G, 1e-5, relu)
    y.backward(gy)
    xd, gd, bd = x.double().clone().requires_grad_(True), gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    yd = torch.nn.functional.group_norm(xd.t().unsqueeze(0), G, gd, bd, 1e-5).squeeze(0).t()
    if relu:
        yd = torch.relu(yd)
    yd.backward(gy.double())
    assert float((y.double() - yd).abs().max()) < 1e-4
    for got, want in ((xa.grad, xd.grad), (ga.grad, gd.grad), (ba.grad, bd.grad)):
        scale = float(want.abs().max()) + 1e-12
        assert float((got.double() - want).abs().max()) / scale < 1e-4


@pytest.mark.parametrize("no_deform", [0, 1])
def test_slice_head_matches_float64(dev, no_deform):
    """csrc/ltn_slice_head.cu (the per-point tail of SliceFastCUDALatticeModule, models.py:232,465, fused into two kernels)
    against the same recipe spelled out in float64 torch: gather -> minus gamma*max+beta -> Linear -> GroupNorm(18) + ReLU ->
    Linear = delta weights -> slice of the class scores -> bias -> log-softmax.  Absent vertices (id -1), a device-side point
    count below the capacity (static-capacity mode) and K = 26 of 32 score columns are all exercised.  Tolerance 2e-5 of
    abs-max (36-term fp32 dot products; the GroupNorm statistics are double precision in both)."""
    from temporal_latticenet_b200 import _lib
    g = torch.Generator().manual_seed(5 + no_deform)
    cap, N, V, K = 6000, 5317, 800, 26
    bott = torch.randn(V, 8, generator=g)
    scores = torch.randn(V, 32, generator=g)
    idx = torch.randint(0, V, (4 * cap,), generator=g).int()
    idx[torch.rand(4 * cap, generator=g) < 0.05] = -1
    w = torch.rand(4 * cap, generator=g)
    gamma, beta = torch.rand(9, generator=g) + 0.5, 0.1 * torch.randn(9, generator=g)
    W1, W2, b2 = torch.randn(36, 36, generator=g) / 6, 0.1 * torch.randn(4, 36, generator=g), 0.1 * torch.randn(4, generator=g)
    gw, gb, cb = torch.rand(36, generator=g) + 0.5, 0.1 * torch.randn(36, generator=g), torch.randn(K, generator=g)

    d = lambda t: t.to(dev).contiguous()   # noqa: E731
    n_dev = torch.tensor([N], dtype=torch.int32, device=dev)
    logits = torch.full((cap, K), float("nan"), device=dev)
    logsm = torch.full((cap, K), float("nan"), device=dev)
    sums = torch.empty(18, 2, dtype=torch.float64, device=dev)
    t = [d(x) for x in (bott, scores, idx, w, gamma, beta, W1, gw, gb, W2, b2, cb)]
    p = _lib.ptr
    rc = _lib.load().ltn_slice_head(p(t[0]), V, None, p(t[1]), 32, p(t[2]), p(t[3]), cap, p(n_dev), p(t[4]), p(t[5]), p(t[6]), p(t[7]),
                                    p(t[8]), 1e-5, p(t[9]), p(t[10]), p(t[11]), K, no_deform, p(sums), p(logits), p(logsm), _lib.stream())
    assert rc == 0
    torch.cuda.synchronize()

    i64, ok = idx[: 4 * N].long().clamp(min=0), (idx[: 4 * N] >= 0)
    wv = torch.where(ok, w[: 4 * N], torch.zeros(())).double()
    rows = torch.cat([wv[:, None] * bott.double()[i64], wv[:, None]], 1).view(N, 4, 9)
    mx = rows.max(1, keepdim=True)[0]
    gg = (rows - (gamma.double() * mx + beta.double())).reshape(N, 36)
    h = gg @ W1.double().t()
    y = torch.relu(torch.nn.functional.group_norm(h.t().unsqueeze(0), 18, gw.double(), gb.double(), 1e-5).squeeze(0).t())
    dw = y @ W2.double().t() + b2.double()
    ww = w[: 4 * N].double() + (0.0 if no_deform else 1.0) * dw.reshape(-1)
    ww = torch.where(ok, ww, torch.zeros((), dtype=torch.float64))
    want = (ww[:, None] * scores.double()[i64][:, :K]).view(N, 4, K).sum(1) + cb.double()
    got = logits[:N].cpu().double()
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max())
    want_ls = torch.log_softmax(want, 1)
    assert float((logsm[:N].cpu().double() - want_ls).abs().max()) < 2e-5 * float(want_ls.abs().max())
    assert torch.isnan(logits[N:]).all() and torch.isnan(logsm[N:]).all()   # rows beyond the live count stay untouched
