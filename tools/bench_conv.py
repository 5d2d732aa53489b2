"""Micro-benchmark of the fused tensor-core convolution against im2row + cuBLAS fp32 on the shapes of the model
(20 launches inside a CUDA graph, CUDA events around the replay)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, funcs, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
ls = Lattice(100000, 0.6, device=dev)
ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
V = ls.nr_lattice_vertices(); nbr = ls.neighbours()
c1 = ls.create_coarse_verts(); V1 = c1.nr_lattice_vertices(); nbr1 = c1.neighbours()
nbr_fin = ls.neighbours(c1, mode=2)
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
print("V0", V, "V1", V1)
cases = [("conv V0 64->64", V, V, nbr, 64, 64), ("conv V0 128->64", V, V, nbr, 128, 64), ("conv V0 192->192", V, V, nbr, 192, 192),
         ("conv V1 128->128", V1, V1, nbr1, 128, 128), ("finefy V1->V0 256->128", V, V1, nbr_fin, 256, 128),
         ("1x1 V0 192->192", V, V, None, 192, 192), ("gru gates V0 192->576", V, V, None, 192, 576)]
if len(sys.argv) > 1: cases = cases[:3]
for name, Vq, Vx, nb, C, F in cases:
    x = torch.randn(Vx, C, device=dev)
    S = 9 if nb is not None else 1
    W = torch.randn(S * C, F, device=dev) / (S * C) ** 0.5
    wt = ops.k_major(W)
    out = torch.empty(Vq, F, device=dev)
    fl = 2.0 * Vq * S * C * F
    t3 = graph_time(lambda: ops.conv_tc(x, nb, wt, nr_rows=Vq, out=out))
    t1 = graph_time(lambda: ops.conv_tc(x, nb, wt, nr_rows=Vq, out=out, passes=1))
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    t16 = graph_time(lambda: ops.conv_tc(x, nb, wt, nr_rows=Vq, out=out, operands="f16", flag=flag)) if C % 64 == 0 else float("nan")
    rows = torch.empty(Vq, S * C, device=dev)
    if nb is not None:
        tb = graph_time(lambda: torch.mm(funcs.im2row_raw(x, nb), W, out=out))
    else:
        tb = graph_time(lambda: torch.mm(x, W, out=out))
    print("%-26s f16x3 %.1f us %6.1f TF/s | tf32x3 %.1f us %6.1f TF/s | tf32x1 %.1f us %6.1f TF/s | im2row+cuBLAS fp32 %.1f us %6.1f TF/s" % (name, 1e3*t16, fl / t16 / 1e9, 1e3*t3, fl / t3 / 1e9, 1e3*t1, fl / t1 / 1e9, 1e3*tb, fl / tb / 1e9))
