"""Host-side profile of the end-to-end path (submit / collect with pinned host buffers): cProfile top entries.
   python tools/prof_e2e.py [lanes] [lockstep|streams]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as G
G.build()
from temporal_latticenet_b200.engine import LockstepRunner, MultiWindowRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
wins = bench.make_windows(4, 1000)
host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in wins]
devw = [[(p.to(dev), v.to(dev)) for p, v in w] for w in host]
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Runner = MultiWindowRunner if (len(sys.argv) > 2 and sys.argv[2] == "streams") else LockstepRunner
multi = Runner(bench.CFG, 26, dev, lanes=lanes).prepare(devw[0], seeded_state, devw)
def dev_loop(steps):
    for i in range(0, steps, lanes):
        multi.infer_windows_device([devw[(i + j) % len(devw)] for j in range(lanes)])
    torch.cuda.synchronize()
dev_loop(8)
t0 = time.perf_counter(); dev_loop(24); dt = time.perf_counter() - t0
print("device-resident %.1f scans/s (%.2f ms/window)" % (4 * 24 / dt, 1e3 * dt / 24))
t0 = time.perf_counter()
for i in range(0, 24, lanes):
    multi.infer_windows_device([devw[(i + j) % len(devw)] for j in range(lanes)])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time to QUEUE 24 windows %.2f ms, then %.2f ms until the device was done" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
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
