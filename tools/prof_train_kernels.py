"""Device-side profile of the training step (config 4): torch.profiler (CUPTI) over 2 steps after warm-up, kernels
aggregated by name and sorted by device time.   python tools/prof_train_kernels.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from temporal_latticenet_b200 import synthetic
from temporal_latticenet_b200.seeding import seeded_state
from temporal_latticenet_b200.train import WindowTrainer
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
wins = bench.make_windows(1, 1000)
devw = [[(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w] for w in wins]
tgt = torch.from_numpy(synthetic.labels(wins[0][-1][0].shape[0], 26, seed=0)).to(dev)
tr = WindowTrainer(bench.CFG, 26, dev)
tr.materialise(devw[0], tgt, seeded_state)
for _ in range(3):
    tr.step(devw[0], tgt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(devw[0], tgt)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0 and e.device_type.name == "CUDA"]
if not ev:
    ev = [e for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
tot = sum(e.device_time_total for e in ev)
print("2 training steps: %d kernel names, %.1f ms of device time per step" % (len(ev), tot / 2e3))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:45]:
    print("%6d %10.1f us/step %5.1f%%  avg %8.1f us  %s" % (e.count // 2, e.device_time_total / 2, 100 * e.device_time_total / tot,
                                                            e.device_time_total / max(e.count, 1), e.key[:110]))
