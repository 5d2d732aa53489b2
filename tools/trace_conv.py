"""Phase breakdown of the fused convolution kernel from its in-kernel globaltimer stamps (ltn_conv_trace):
python tools/trace_conv.py  -> per shape: median over CTAs of the time spent in each phase, and the launch-to-launch gap."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic, _lib
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
ls = Lattice(100000, 0.6, device=dev)
ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
V = ls.nr_lattice_vertices(); nbr = ls.neighbours()
lib = _lib.load()
names = ["setup", "fill", "mainloop", "drain", "epilogue", "teardown"]
for name, nb, C, F, mode in [("conv 64->64 f16", nbr, 64, 64, "f16"), ("conv 192->192 f16", nbr, 192, 192, "f16"), ("conv 192->192 tf32", nbr, 192, 192, "tf32"),
                             ("1x1 192->192 f16", None, 192, 192, "f16"), ("gates 192->576 f16", None, 192, 576, "f16")]:
    S = 9 if nb is not None else 1
    x = torch.randn(V, C, device=dev); W = torch.randn(S * C, F, device=dev) / (S * C) ** 0.5
    wt = ops.k_major(W); out = torch.empty(V, F, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    run = lambda: ops.conv_tc(x, nb, wt, out=out, operands=mode, flag=flag)
    for _ in range(3): run()
    reps = 6
    bufs = [torch.zeros(1024 * 8, dtype=torch.int64, device=dev) for _ in range(reps)]
    torch.cuda.synchronize()
    for b in bufs:
        lib.ltn_conv_trace(_lib.ptr(b)); run()
    lib.ltn_conv_trace(None)
    torch.cuda.synchronize()
    T = [b.cpu().numpy().reshape(-1, 8) for b in bufs]
    T = [t[t[:, 0] > 0] for t in T]
    t = T[-1].astype(np.float64)
    d = np.diff(t[:, :7], axis=1) / 1e3
    span = (t[:, 6].max() - t[:, 0].min()) / 1e3
    gap = (T[-1][:, 0].min() - T[-2][:, 6].max()) / 1e3
    print("%-20s CTAs %4d | kernel span %.1f us, gap after previous launch %.1f us | entry spread %.1f us | median phase us: %s" % (
        name, len(t), span, gap, (t[:, 0].max() - t[:, 0].min()) / 1e3, "  ".join("%s %.1f" % (n, np.median(d[:, i])) for i, n in enumerate(names))))
