"""Per-parameter gradient difference between the fused training path (funcs._FusedConv) and the unfused autograd composition
(LTN_TRAIN_UNFUSED=1) on the same 2-frame window; prints them in module order.   python tools/debug_train_fused.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from temporal_latticenet_b200 import synthetic
from temporal_latticenet_b200.config import ConfigParser
from temporal_latticenet_b200.lattice import Lattice, ModelParams
from temporal_latticenet_b200.lovasz import LovaszSoftmax
from temporal_latticenet_b200.model import LatticeNetSeq
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
cfg = bench.CFG
frames = []
for fp, fv in synthetic.window(2, frames=2, nr_points=20000):
    keep = np.linalg.norm(fp[:, [0, 2]], axis=1) < 10.0
    frames.append((torch.from_numpy(np.ascontiguousarray(fp[keep][:6000])).to(dev), torch.from_numpy(np.ascontiguousarray(fv[keep][:6000])).to(dev)))
target = torch.from_numpy(np.random.default_rng(0).integers(0, 26, frames[-1][0].shape[0])).to(dev)
model = LatticeNetSeq(26, ModelParams.create(cfg), ConfigParser(cfg)).to(dev)
model.train(True)
lov, nll = LovaszSoftmax(ignore_index=0), torch.nn.NLLLoss(ignore_index=0)
def window_loss():
    model.reset_sequence()
    ls = Lattice.create(cfg, "lattice")
    for i, (p, v) in enumerate(frames):
        out, _, ls = model(ls, p, v, i != len(frames) - 1, True)
    return 0.5 * lov(out, target) + 0.5 * nll(out, target)
with torch.no_grad():
    window_loss()
model.load_state_dict(seeded_state({k: tuple(v.shape) for k, v in model.state_dict().items()}))
res = {}
for mode in ("1", "0"):
    os.environ["LTN_TRAIN_UNFUSED"] = mode
    model.zero_grad(set_to_none=True)
    loss = window_loss()
    loss.backward()
    res[mode] = (float(loss), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
print("loss unfused %.7f fused %.7f" % (res["1"][0], res["0"][0]))
for k, g1 in res["1"][1].items():
    g0 = res["0"][1].get(k)
    if g0 is None:
        print("%-70s MISSING in fused" % k); continue
    e = float((g0 - g1).abs().max()) / (float(g1.abs().max()) + 1e-30)
    print("%-70s %.2e %s" % (k, e, "<<<" if e > 1e-3 else ""))
