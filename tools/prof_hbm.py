"""The HBM-side kernels the metric names (im2row, slice, splat, distribute) on the accumulated 4-scan cloud, one launch
each, for `ncu --set full -k regex:"k_im2row|k_slice|k_splat|k_distribute_rows|k_insert_points"` (profiles/README.md)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import _lib, synthetic  # noqa: E402
from temporal_latticenet_b200.lattice import Lattice  # noqa: E402

dev = torch.device("cuda:0")
win = synthetic.window(1000, frames=4)
pos = torch.from_numpy(np.concatenate([f[0] for f in win], 0)).to(dev)
val = torch.from_numpy(np.concatenate([f[1] for f in win], 0)).to(dev)
N = pos.shape[0]
lib, p = _lib.load(), _lib.ptr
for rep in range(2):   # second pass = warm caches / allocator; ncu picks launches by -s/-c
    ls = Lattice(100000, 0.6, device=dev)
    rows, idx, w = ls.distribute(pos, val, True)
    V = ls.nr_lattice_vertices()
    nbr = ls.neighbours()
    feat = torch.randn(V, 192, device=dev)
    out = torch.empty(V, 9 * 192, device=dev)
    lib.ltn_im2row(p(feat), V, None, p(nbr), V, None, 192, p(out), _lib.stream())
    vals = torch.randn(V, 32, device=dev)
    sl = torch.empty(N, 32, device=dev)
    lib.ltn_slice(p(vals), V, 32, p(idx), p(w), N, p(sl), _lib.stream())
    acc = torch.zeros(V, 2, device=dev)
    lib.ltn_splat(p(val), N, 1, p(idx), p(w), p(acc), V, _lib.stream())
torch.cuda.synchronize()
print("ok N=%d V=%d" % (N, V))
