import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from temporal_latticenet_b200 import ops, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
ls = Lattice(100000, 0.6, device=dev)
ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
V = ls.nr_lattice_vertices(); nbr = ls.neighbours()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for C, F in [(192, 192), (64, 64)]:
    x = torch.randn(V, C, device=dev); W = torch.randn(9 * C, F, device=dev) / (9 * C) ** 0.5; wt = ops.k_major(W)
    for passes in (3, 1):
        print(C, F, "passes", passes, "dbg", os.environ.get("LTN_CONV_DBG"), "%.1f us" % (1e3 * timeit(lambda: ops.conv_tc(x, nbr, wt, passes=passes))))
