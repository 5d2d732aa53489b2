"""Micro-benchmark: the same layer of 4 windows as ONE batched persistent launch (csrc/ltn_conv_batched.cu) against 4
single launches of k_conv_tc (20 repetitions inside a CUDA graph, CUDA events around the replay, one stream).
Row counts = the V0 / V1 / V2 sizes of the 4 frames of the timed window (seed 1000)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic, _lib
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
win = synthetic.window(1000, frames=4)
ls = Lattice(100000, 0.6, device=dev)
levels = []
for f, (p, v) in enumerate(win):
    ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), f == 0)
    c1 = ls.create_coarse_verts(); c2 = c1.create_coarse_verts()
    levels.append((ls.nr_lattice_vertices(), c1.nr_lattice_vertices(), c2.nr_lattice_vertices()))
nbrs = (ls.neighbours(), c1.neighbours(), c2.neighbours())
print("vertex counts per frame (V0, V1, V2):", levels)
mode = sys.argv[1] if len(sys.argv) > 1 else "frames"   # frames: the 4 lanes are at frames 0..3 sizes;  last: all at the last frame's size
def graph_time(fn, n=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
class Col:
    def __init__(self): self.reqs = []
    def request(self, d): self.reqs.append(d)
cases = [("conv V0 64->64", 0, 64, 64, 9), ("conv V0 128->64", 0, 128, 64, 9), ("conv V0 192->192", 0, 192, 192, 9),
         ("conv V1 128->128", 1, 128, 128, 9), ("conv V2 64->64", 2, 64, 64, 9), ("1x1 V0 192->192", 0, 192, 192, 1),
         ("gates V0 192->576", 0, 192, 576, 1), ("gates V0 128->384", 0, 128, 384, 1), ("1x1 V2 64->256", 2, 64, 256, 1), ("1x1 V2 256->64", 2, 256, 64, 1)]
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for name, lvl, C, F, S in cases:
    Vs = [levels[f][lvl] for f in range(4)] if mode == "frames" else [levels[3][lvl]] * 4
    W = torch.randn(S * C, F, device=dev) / (S * C) ** 0.5
    wt = ops.k_major(W)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    xs = [torch.randn(Vb, C, device=dev) for Vb in Vs]
    gns = [(ops.gn_sums(x, ops.gn_groups(C)), gamma, beta, 1e-5) for x in xs] if C <= 256 else [None] * 4
    outs = [torch.empty(Vb, F, device=dev) for Vb in Vs]
    sums = [torch.zeros(ops.gn_groups(F), 2, dtype=torch.float64, device=dev) if F % 32 == 0 else None for _ in Vs]
    nb = [nbrs[lvl][:Vb].contiguous() if S == 9 else None for Vb in Vs]
    def single():
        for b in range(4):
            ops.conv_tc(xs[b], nb[b], wt, nr_rows=Vs[b], gn=gns[b], relu=True, out=outs[b], out_sums=sums[b], operands="f16", flag=flag)
    col = Col(); ops._BATCH.ctx = col; single(); ops._BATCH.ctx = None
    fl = sum(2.0 * Vb * S * C * F for Vb in Vs)
    ts = graph_time(single)
    tb = graph_time(lambda: ops.conv_tc_batched(col.reqs))
    tiles = sum((Vb + 127) // 128 for Vb in Vs)
    print("%-20s rows %-28s tiles %4d | 4 single launches %7.1f us %6.1f TF/s | 1 batched %7.1f us %6.1f TF/s | x%.2f" % (
        name, Vs, tiles, 1e3 * ts, fl / ts / 1e9, 1e3 * tb, fl / tb / 1e9, ts / tb))
