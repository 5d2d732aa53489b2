"""Host-side profile of the end-to-end path (MultiWindowRunner.submit / collect with pinned host buffers): cProfile top entries."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as G
G.build()
from temporal_latticenet_b200.engine import MultiWindowRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
wins = bench.make_windows(2, 1000)
host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in wins]
devw = [[(p.to(dev), v.to(dev)) for p, v in w] for w in host]
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
multi = MultiWindowRunner(bench.CFG, bench.NR_CLASSES, dev, lanes=lanes).prepare(devw[0], seeded_state, devw)
def loop(steps):
    pending = None
    for i in range(0, steps, lanes):
        t = multi.submit([host[(i + j) % len(host)] for j in range(lanes)])
        if pending is not None:
            multi.collect(pending)
        pending = t
    multi.collect(pending)
    torch.cuda.synchronize()
loop(6)
t0 = time.perf_counter(); loop(24); dt = time.perf_counter() - t0
print("e2e %.1f scans/s (%.2f ms/window)" % (4 * 24 / dt, 1e3 * dt / 24))
# host-only cost of submit (device work queued, not waited for)
t0 = time.perf_counter(); tk = multi.submit([host[j % len(host)] for j in range(lanes)]); t1 = time.perf_counter(); multi.collect(tk); t2 = time.perf_counter()
print("submit host time %.2f ms for %d windows, collect (incl. wait) %.2f ms" % (1e3 * (t1 - t0), lanes, 1e3 * (t2 - t1)))
pr = cProfile.Profile(); pr.enable(); loop(24); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
