"""Timing of the slice head (ltn_slice_head: statistics + apply kernels) on one scan of the timed window, back to back on one
stream (CUDA events).  LTN_HEAD_STATS_BPS=<blocks per SM of the statistics kernel> is read once per process.
python tools/bench_slice_head.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import _lib, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
win = synthetic.window(1000, frames=4)
ls = Lattice(100000, 0.6, device=dev)
for f, (p_, v_) in enumerate(win):
    rows, idx, w = ls.distribute(torch.from_numpy(p_).to(dev), torch.from_numpy(v_).to(dev), f == 0)
V, N, K = ls.nr_lattice_vertices(), win[-1][0].shape[0], 26
g = torch.Generator().manual_seed(0)
d = lambda *s: torch.randn(*s, generator=g).to(dev)
bott, scores = d(V, 8), d(V, 32)
a = [d(9).abs() + 0.5, d(9) * 0.1, d(36, 36) / 6, d(36).abs() + 0.5, d(36) * 0.1, d(4, 36) * 0.1, d(4) * 0.1, d(K)]
logits, logsm = torch.empty(N, K, device=dev), torch.empty(N, K, device=dev)
sums = torch.empty(18, 2, dtype=torch.float64, device=dev)
p = _lib.ptr
def run():
    rc = _lib.load().ltn_slice_head(p(bott), V, None, p(scores), 32, p(idx), p(w), N, None, p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), 1e-5,
                                    p(a[5]), p(a[6]), p(a[7]), K, 0, p(sums), p(logits), p(logsm), _lib.stream())
    assert rc == 0
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
print("LTN_HEAD_STATS_BPS=%s: %.1f us per call (N=%d V=%d), checksum %.6f" % (os.environ.get("LTN_HEAD_STATS_BPS", "default"), 1e3 * e0.elapsed_time(e1) / 50, N, V, float(logsm.double().sum())))
