"""One window frame's PointNet front end (distribute + fused MLP / segmented max) for `ncu --set full -k regex:k_pointnet`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporal_latticenet_b200 import _lib, synthetic
from temporal_latticenet_b200.lattice import Lattice
dev = torch.device("cuda:0")
p, v = synthetic.window(1000, frames=1)[0]
lib, P = _lib.load(), _lib.ptr
g = torch.Generator().manual_seed(0)
w1, b1 = torch.randn(16, 4, generator=g).to(dev), torch.randn(16, generator=g).to(dev)
w2, b2 = (torch.randn(32, 16, generator=g) / 4).to(dev), torch.randn(32, generator=g).to(dev)
w3, b3 = (torch.randn(64, 32, generator=g) / 6).to(dev), torch.randn(64, generator=g).to(dev)
for rep in range(3):
    ls = Lattice(100000, 0.6, device=dev)
    rows, idx, w = ls.distribute(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), True)
    V = ls.nr_lattice_vertices()
    packed = torch.empty(V, 64, dtype=torch.int64, device=dev)
    out = torch.empty(V, 128, device=dev)
    rc = lib.ltn_pointnet(P(rows), rows.shape[1], P(idx), rows.shape[0], None, P(w1), P(b1), P(w2), P(b2), P(w3), P(b3), V, None,
                          P(packed), P(ls.vert_acc()) if hasattr(ls, "vert_acc") else P(ls._vert_acc), 4, P(out), _lib.stream())
    assert rc == 0
torch.cuda.synchronize()
print("ok R=%d V=%d" % (rows.shape[0], V), float(out.abs().mean()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.ltn_pointnet(P(rows), rows.shape[1], P(idx), rows.shape[0], None, P(w1), P(b1), P(w2), P(b2), P(w3), P(b3), V, None,
                     P(packed), P(ls._vert_acc), 4, P(out), _lib.stream())
e1.record(); torch.cuda.synchronize()
print("ltn_pointnet (memset + mlp_max + decode): %.1f us per call" % (1e3 * e0.elapsed_time(e1) / 10))
import ctypes, numpy as np
w12 = np.ascontiguousarray(np.concatenate([t.cpu().numpy().reshape(-1) for t in (w1, b1, w2, b2)]).astype(np.float32))
W12 = w12.ctypes.data_as(ctypes.c_void_p)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(3):
    lib.ltn_pointnet_tc(P(rows), rows.shape[1], P(idx), rows.shape[0], None, W12, P(w3), P(b3), V, None,
                        P(packed), P(ls._vert_acc), 4, P(out), 5, P(flag), _lib.stream())
e0.record()
for _ in range(10):
    lib.ltn_pointnet_tc(P(rows), rows.shape[1], P(idx), rows.shape[0], None, W12, P(w3), P(b3), V, None,
                        P(packed), P(ls._vert_acc), 4, P(out), 5, P(flag), _lib.stream())
e1.record(); torch.cuda.synchronize()
print("ltn_pointnet_tc (memset + tc + decode): %.1f us per call, flag %d" % (1e3 * e0.elapsed_time(e1) / 10, int(flag.item())))
tr = torch.zeros(16, dtype=torch.int64, device=dev)
lib.ltn_pointnet_trace(P(tr))
lib.ltn_pointnet_tc(P(rows), rows.shape[1], P(idx), rows.shape[0], None, W12, P(w3), P(b3), V, None,
                    P(packed), P(ls._vert_acc), 4, P(out), 5, P(flag), _lib.stream())
lib.ltn_pointnet_trace(None)
t = tr.cpu().numpy()
names = ["load+MLP+tmem st", "sync, hash insert, compaction", "table init, MMA wait, tmem ld", "sync + atomicMax", "sync + atomicMin", "sync + flush", "sync"]
print("one tile of k_pointnet_tc (cycles):", ", ".join("%s %d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), "| total", t[7] - t[0], "| distinct vertices", t[8])
