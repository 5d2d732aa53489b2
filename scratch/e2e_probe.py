import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
from temporal_latticenet_b200.runner import WindowRunner
from temporal_latticenet_b200.seeding import seeded_state
dev = torch.device("cuda:0")
w = bench.make_windows(1, 1000)
host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in x] for x in w]
devw = [[(p.to(dev), v.to(dev)) for p, v in x] for x in host]
r = WindowRunner(bench.CFG, 26, dev)
r.materialise_parameters(devw[0], seeded_state)
for i in range(3): r.infer_window_device(devw[0])
torch.cuda.synchronize()
for name, fn in (("device", lambda: r.infer_window_device(devw[0])), ("host", lambda: r.infer_window(host[0])), ("device", lambda: r.infer_window_device(devw[0]))):
    ts = []
    for i in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print(name, ["%.1f" % t for t in ts])
# phases of the host path
t0 = time.perf_counter(); frames = [(p.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for p, v in host[0]]; torch.cuda.synchronize(); t1 = time.perf_counter()
out = r.infer_window_device(frames); torch.cuda.synchronize(); t2 = time.perf_counter()
lab = out.argmax(1); torch.cuda.synchronize(); t3 = time.perf_counter()
h = torch.empty(lab.shape[0], dtype=torch.int64).pin_memory(); t4 = time.perf_counter()
h.copy_(lab, non_blocking=True); torch.cuda.synchronize(); t5 = time.perf_counter()
print("h2d %.2f fwd %.2f argmax %.2f pin %.2f d2h %.2f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3): r.infer_window_device(devw[0])
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
