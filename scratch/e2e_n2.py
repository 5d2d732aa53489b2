import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import torch.distributed as dist
import bench
from temporal_latticenet_b200.engine import MultiWindowRunner
from temporal_latticenet_b200.seeding import seeded_state
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
w = bench.make_windows(2, 1000 + 100 * rank)
host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in x] for x in w]
devw = [[(p.to(dev), v.to(dev)) for p, v in x] for x in host]
multi = MultiWindowRunner(bench.CFG, 26, dev, lanes=3).prepare(devw[0], seeded_state, devw)
for rep in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ts = []; tc = []
    pending = None
    for i in range(0, 24, 3):
        a = time.perf_counter()
        ticket = multi.submit([host[(i + j) % 2] for j in range(3)])
        b = time.perf_counter()
        if pending is not None: multi.collect(pending)
        c = time.perf_counter()
        ts.append(1e3 * (b - a)); tc.append(1e3 * (c - b))
        pending = ticket
    multi.collect(pending); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("rank", rank, "rep", rep, "e2e scans/s %.0f" % (96 / dt), "submit ms", ["%.1f" % x for x in ts], "collect ms", ["%.1f" % x for x in tc], flush=True)
    # device-resident for comparison
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(0, 24, 3): multi.infer_windows_device([devw[(i + j) % 2] for j in range(3)])
    torch.cuda.synchronize(); print("rank", rank, "device scans/s %.0f" % (96 / (time.perf_counter() - t0)), flush=True)
if world > 1: dist.destroy_process_group()
