"""ncu targets: the two slice-head kernels and the tensor-core weight gradient on shapes of the timed window.
python tools/prof_slice_head.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import _lib, funcs, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
win = synthetic.window(1000, frames=4)
ls = Lattice(100000, 0.6, device=dev)
for f, (p, v) in enumerate(win):
    rows, idx, w = ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), f == 0)
V, N, K = ls.nr_lattice_vertices(), win[-1][0].shape[0], 26
g = torch.Generator().manual_seed(0)
d = lambda *s: torch.randn(*s, generator=g).to(dev)
bott, scores = d(V, 8), d(V, 32)
args = [d(9).abs() + 0.5, d(9) * 0.1, d(36, 36) / 6, d(36).abs() + 0.5, d(36) * 0.1, d(4, 36) * 0.1, d(4) * 0.1, d(K)]
logits, logsm = torch.empty(N, K, device=dev), torch.empty(N, K, device=dev)
sums = torch.empty(18, 2, dtype=torch.float64, device=dev)
p = _lib.ptr
for _ in range(3):
    rc = _lib.load().ltn_slice_head(p(bott), V, None, p(scores), 32, p(idx), p(w), N, None, p(args[0]), p(args[1]), p(args[2]), p(args[3]),
                                    p(args[4]), 1e-5, p(args[5]), p(args[6]), p(args[7]), K, 0, p(sums), p(logits), p(logsm), _lib.stream())
    assert rc == 0
act, dy = torch.relu(d(V, 192)), d(V, 192) * 1e-3
for _ in range(3):
    funcs.conv_bwd_weight(act, dy, ls.neighbours())
torch.cuda.synchronize()
print("ok N=%d V=%d: slice head moves %.1f MB (idx, w, logits, log-softmax), bwd_weight 192->192 S=9: %.2f GFLOP" % (
    N, V, (N * 32 + 2 * N * K * 4) / 1e6, 2.0 * V * 9 * 192 * 192 / 1e9))
