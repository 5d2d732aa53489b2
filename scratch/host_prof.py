import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
from temporal_latticenet_b200.runner import WindowRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
w = bench.make_windows(1, 1000)
devw = [[(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in x] for x in w]
r = WindowRunner(bench.CFG, 26, dev)
r.materialise_parameters(devw[0], seeded_state)
for i in range(3): r.infer_window_device(devw[0])
torch.cuda.synchronize()
ts = []
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r.infer_window_device(devw[0]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
print("host-enqueue ms, total ms:", ["%.1f/%.1f" % t for t in ts])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3): r.infer_window_device(devw[0])
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
