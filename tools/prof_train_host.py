"""Host-side profile of the training step (config 4): cProfile over 3 steps after warm-up, top functions by cumulative and
own time, plus device time of a step from CUDA events.   python tools/prof_train_host.py"""
import cProfile, io, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from temporal_latticenet_b200 import synthetic
from temporal_latticenet_b200.seeding import seeded_state
from temporal_latticenet_b200.train import WindowTrainer
dev = torch.device("cuda:0")
wins = bench.make_windows(1, 1000)
devw = [[(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in w] for w in wins]
tgt = torch.from_numpy(synthetic.labels(wins[0][-1][0].shape[0], 26, seed=0)).to(dev)
tr = WindowTrainer(bench.CFG, 26, dev)
tr.materialise(devw[0], tgt, seeded_state)
for _ in range(3):
    tr.step(devw[0], tgt)
torch.cuda.synchronize()
t0 = time.perf_counter()
tr.step(devw[0], tgt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("one step: host returns after %.1f ms, device done after %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
# forward / backward / optimizer split (host + device, synchronised)
def timed(fn):
    torch.cuda.synchronize(); a = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, 1e3 * (time.perf_counter() - a)
loss, t_f = timed(lambda: tr.forward_window(devw[0], tgt))
tr.optimizer.zero_grad(set_to_none=False); tr.allreduce.prepare()
_, t_b = timed(lambda: loss.backward())
_, t_o = timed(lambda: (tr.allreduce(), tr.optimizer.step()))
print("forward %.1f ms, backward %.1f ms, all-reduce + optimizer %.1f ms" % (t_f, t_b, t_o))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    tr.step(devw[0], tgt)
torch.cuda.synchronize()
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(35)
    print(s.getvalue()[:6000])
