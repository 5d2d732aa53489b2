"""Which layers still need a separate GroupNorm statistics kernel (k_gn_stats) because no producing epilogue left the sums
behind: python tools/trace_gn_stats.py  (round 1: 3 per full frame -- after the middle GRU, the up-path concatenation, after
the late GRU)."""
import sys, os, collections, traceback
sys.path.insert(0, os.getcwd())
import torch, bench
import __graft_entry__ as G
G.build()
from temporal_latticenet_b200 import ops
from temporal_latticenet_b200.runner import WindowRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
win = bench.make_windows(1, 1000)[0]
fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in win]
r = WindowRunner(bench.CFG, 26, dev).materialise_parameters(fd, seeded_state)
calls = collections.Counter()
orig = ops.gn_sums
def traced(x, groups):
    st = traceback.extract_stack(limit=8)
    key = " <- ".join("%s:%d" % (os.path.basename(f.filename), f.lineno) for f in reversed(st[:-1]) if "temporal_latticenet_b200" in f.filename)[:160]
    calls[(tuple(x.shape), groups, key)] += 1
    return orig(x, groups)
ops.gn_sums = traced
r.infer_window_device(fd)
for k, v in sorted(calls.items(), key=lambda kv: -kv[1]): print(v, k)
