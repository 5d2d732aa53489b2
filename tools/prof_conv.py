"""One representative launch of the dominant kernel for `ncu --set full` (profiles/README.md): the 192->192 lattice
convolution on the V0 lattice of one synthetic scan."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
ls = Lattice(100000, 0.6, device=dev)
ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
V = ls.nr_lattice_vertices(); nbr = ls.neighbours()
C = F = 192
x = torch.randn(V, C, device=dev)
W = torch.randn(9 * C, F, device=dev) / (9 * C) ** 0.5
wt = ops.k_major(W)
passes = 3
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(4):
    out = ops.conv_tc(x, nbr, wt, passes=passes, operands=mode, flag=flag)
torch.cuda.synchronize()
print("ok", V, float(out.abs().mean()))
