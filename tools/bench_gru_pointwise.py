import os, sys, torch
sys.path.insert(0, "/root/repo")
from temporal_latticenet_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.load(); p = _lib.ptr
for C in (64, 128, 192):
    V = 23000
    gi, gh = torch.randn(V, 3 * C, device=dev), torch.randn(V, 3 * C, device=dev)
    h, b = torch.randn(V, C, device=dev), torch.randn(3 * C, device=dev)
    out = torch.empty(V, C, device=dev); sums = torch.zeros(32, 2, dtype=torch.float64, device=dev)
    def t(fn, n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return 1e3 * e0.elapsed_time(e1) / n
    a = t(lambda: lib.ltn_gru_pointwise(p(gi), p(gh), p(h), p(b), V, V - 3000, None, None, C, p(out), _lib.stream()))
    ref = out.clone()
    s = t(lambda: lib.ltn_gru_pointwise_stats(p(gi), p(gh), p(h), p(b), V, V - 3000, None, None, C, p(out), p(sums), 32, _lib.stream()))
    sums.zero_(); lib.ltn_gru_pointwise_stats(p(gi), p(gh), p(h), p(b), V, V - 3000, None, None, C, p(out), p(sums), 32, _lib.stream()); torch.cuda.synchronize()
    want = torch.stack([out.double().view(V, 32, -1).sum((0, 2)), (out.double() ** 2).view(V, 32, -1).sum((0, 2))], 1)
    print("C=%d plain %.1f us, with stats %.1f us, out equal %s, stats rel err %.1e" % (C, a, s, bool(torch.equal(ref, out)), float(((sums - want).abs() / want.abs().clamp(min=1e-9)).max())))
