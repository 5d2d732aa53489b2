"""Timeline of ONE lock-step group inside its replayed graphs: every batched tensor-core launch (in-kernel %globaltimer stamps)
with the GAP before it, i.e. the time the other kernels of the frame (hash build, PointNet, neighbour tables, gates, AFlow,
slicing -- on the per-window streams) keep the tensor-core stream waiting.   python tools/timeline_group.py [lanes]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from temporal_latticenet_b200.engine import LockstepRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wins = bench.make_windows(2, 1000)
devw = [[(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w] for w in wins]
r = LockstepRunner(bench.CFG, 26, dev, lanes=lanes)
r.prepare(devw[0], seeded_state, devw)
group = [devw[j % 2] for j in range(lanes)]
for _ in range(3):
    r.infer_windows_device(group)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r.infer_windows_device(group); e1.record(); torch.cuda.synchronize()
print("group of %d windows: %.3f ms" % (lanes, e0.elapsed_time(e1)))
rec = r.trace_group(group)
tot_k = tot_g = 0.0
for f in range(4):
    rs = [x for x in rec if x["frame"] == f]
    if not rs:
        continue
    prev = None
    fk = fg = 0.0
    print("frame %d: %d batched launches" % (f, len(rs)))
    for x in rs:
        gap = (x["t0_ns"] - prev) / 1e3 if prev is not None else 0.0
        prev = x["t1_ns"]
        fk += x["us"]; fg += gap
        print("  C%-3d S%d F%-3d rows %6d tiles %4d  gap %7.1f us  run %7.1f us  %6.1f TF/s" % (
            x["C"], x["S"], x["F"], x["rows"], x["tiles"], gap, x["us"], x["flop"] / x["us"] / 1e6))
    span = (rs[-1]["t1_ns"] - rs[0]["t0_ns"]) / 1e3
    print("  frame %d: first launch -> last exit %.1f us, tensor-core launches %.1f us, gaps between them %.1f us" % (f, span, fk, fg))
    tot_k += fk; tot_g += fg
print("sum over frames: launches %.1f us, gaps %.1f us (the head of each frame before its first launch and the tail after the last are not in the gaps)" % (tot_k, tot_g))
