"""Representative launches of the dominant kernel for `ncu --set full` (profiles/README.md): lattice convolutions on the V0
lattice of one synthetic scan.  python tools/prof_conv.py [f16|tf32] [C] [F]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import ops, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
ls = Lattice(100000, 0.6, device=dev)
ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
V = ls.nr_lattice_vertices(); nbr = ls.neighbours()
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 192
F = int(sys.argv[3]) if len(sys.argv) > 3 else C
flag = torch.zeros(1, dtype=torch.int32, device=dev)
x = torch.randn(V, C, device=dev)
W = torch.randn(9 * C, F, device=dev) / (9 * C) ** 0.5
wt = ops.k_major(W)
for _ in range(4):
    out = ops.conv_tc(x, nbr, wt, operands=mode, flag=flag)
torch.cuda.synchronize()
print("ok", V, C, F, float(out.abs().mean()))
