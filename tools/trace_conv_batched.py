"""Role timeline of CTA 0 of one batched persistent convolution launch (in-kernel %globaltimer stamps, ltn_conv_batched_detail):
when each tile's neighbour slice was staged, when the gather entered / left it, when the MMA issuer had its first operands,
committed the accumulator, when the epilogue took and returned it.   python tools/trace_conv_batched.py [C] [F] [S]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic, _lib
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
for _ in range(3):
    ops.conv_tc_batched(col.reqs)
torch.cuda.synchronize()
buf = torch.zeros(64, 16, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.ltn_conv_batched_detail(_lib.ptr(buf))
ops.conv_tc_batched(col.reqs)
torch.cuda.synchronize()
lib.ltn_conv_batched_detail(None)
t = buf.cpu().numpy()
t0 = t[t > 0].min()
names = ["issue-in", "cons-in", "cons-out", "mma-pre", "mma-acc-ok", "mma-commit", "epi-in", "epi-out", "B-ready", "A-ready", "tma-done", "meta-done"]
print("shape C%d F%d S%d, rows %s, k-blocks per tile %d; times in us since the first stamp (CTA 0)" % (C, F, S, Vs, S * C // 64))
print("tile " + " ".join("%10s" % n for n in names))
for it in range(64):
    if not (t[it] > 0).any():
        break
    print("%4d " % it + " ".join(("%10.2f" % ((t[it, k] - t0) / 1e3)) if t[it, k] > 0 else "%10s" % "-" for k in range(len(names))))
