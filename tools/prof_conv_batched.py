"""Representative launches of the batched persistent convolution for `ncu --set full` (profiles/README.md): the same
layer of 4 windows (the V0 sizes of the 4 frames of the timed window).   python tools/prof_conv_batched.py [C] [F] [S]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
win = synthetic.window(1000, frames=4)
ls = Lattice(100000, 0.6, device=dev)
Vs = []
for f, (p, v) in enumerate(win):
    ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), f == 0)
    Vs.append(ls.nr_lattice_vertices())
nbr = ls.neighbours()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 192
F = int(sys.argv[2]) if len(sys.argv) > 2 else C
S = int(sys.argv[3]) if len(sys.argv) > 3 else 9
flag = torch.zeros(1, dtype=torch.int32, device=dev)
W = torch.randn(S * C, F, device=dev) / (S * C) ** 0.5
wt = ops.k_major(W)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
xs = [torch.randn(Vb, C, device=dev) for Vb in Vs]
class Col:
    def __init__(self): self.reqs = []
    def request(self, d): self.reqs.append(d)
col = Col(); ops._BATCH.ctx = col
for b, Vb in enumerate(Vs):
    ops.conv_tc(xs[b], nbr[:Vb].contiguous() if S == 9 else None, wt, nr_rows=Vb, gn=(ops.gn_sums(xs[b], ops.gn_groups(C)), gamma, beta, 1e-5),
                relu=True, out_sums=torch.zeros(ops.gn_groups(F), 2, dtype=torch.float64, device=dev) if F % 32 == 0 else None,
                operands="f16", flag=flag)
ops._BATCH.ctx = None
for _ in range(4):
    ops.conv_tc_batched(col.reqs)
torch.cuda.synchronize()
flops = sum(2.0 * Vb * S * C * F for Vb in Vs)
bytes_ = sum(Vb * 4 * (C + F) for Vb in Vs) + 4 * S * C * F
print("ok rows", Vs, "C", C, "F", F, "S", S, "algorithmic GFLOP %.3f MB %.2f" % (flops / 1e9, bytes_ / 1e6))
