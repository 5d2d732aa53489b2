#!/usr/bin/env python3
"""bench.py -- 4-frame scans/sec of the Temporal LatticeNet hot path on B200 (BASELINE.json metric).

A "step" is one 4-frame window (BASELINE config 3: rnn_modules [gru,gru,aflow,gru], frames 4, scope 3,
sigma 0.6, 26 classes, inference, seeded weights because the pretrained checkpoint is missing) of
synthetic SemanticKITTI-shaped scans (~125k points each) pushed through the window runner.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU)
  python bench.py --impl reference ...                            CPU arm: the scalar oracle + torch CPU
  python bench.py --workload config5|config5-accumulated|config2  the other BASELINE configs (same contract)
  python bench.py --mode train                                    BASELINE config 4 (training step, NCCL all-reduce)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

CFG = os.path.join(REPO, "configs", "lnn_eval_semantic_kitti.cfg")
METRIC = "4-frame scans/sec"
UNIT = "scans/s"
FRAMES = 4
WORKLOADS = {
    # name: (description, rnn_modules or None = cfg default, nr_classes, frames per window, accumulate)
    "config3": ("config3: LNN_SEQ [gru,gru,aflow,gru], 4 frames, scope 3, sigma 0.6, 26 classes, inference, fp32", None, 26, 4, False),
    "config5": ("config5: LNN_SEQ [aflow,aflow,aflow,aflow], 4 frames, scope 3, sigma 0.6, 26 classes, inference, fp32",
                ["aflow", "aflow", "aflow", "aflow"], 26, 4, False),
    "config5-accumulated": ("config5 (accumulate_clouds): the window's 4 scans as ONE ~500k-point lattice, [aflow x4], 26 classes, inference, fp32",
                            ["aflow", "aflow", "aflow", "aflow"], 26, 1, True),
    "config2": ("config2: single-frame LatticeNet forward (conv/coarsen/finefy/slice_classify), 20 classes, inference, fp32", None, 20, 1, False),
}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock / throttle-reason samples DURING the timed regions.  In-process NVML (the library nvidia-smi itself
    queries) from a background thread every 50 ms: a `nvidia-smi -lms` child process was measured to stall the
    launching thread for tens of milliseconds per query.  Falls back to the nvidia-smi child when the NVML binding is
    unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.samples, self.nvml = index, None, [], [], None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_CLOCKS_MS", "200")], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.h))
                self.samples.append((sm, [name for name, bit in bits if mask & bit]))
            except Exception:
                pass
            self._stop.wait(float(os.environ.get("BENCH_CLOCKS_MS", "50")) / 1e3)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            sm = sorted(s for s, _ in self.samples)
            reasons = sorted({r for _, rs in self.samples for r in rs})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# inputs and configuration (IDENTICAL for both arms: the driver compares the `config` objects)
# ------------------------------------------------------------------------------------------------
def make_windows(nr_windows, seed0, frames=FRAMES, accumulate=False):
    import numpy as np
    from temporal_latticenet_b200 import synthetic
    out = []
    for i in range(nr_windows):
        w = synthetic.window(seed0 + i, frames=FRAMES if accumulate else frames, scope=3)
        if accumulate:   # kitti_dataloader.py:198-201: the window's scans concatenated into one cloud
            w = [(np.ascontiguousarray(np.concatenate([p for p, _ in w], 0)), np.ascontiguousarray(np.concatenate([v for _, v in w], 0)))]
        out.append(w)
    return out


def workload_cfg(name):
    """cfg file of a workload (the KITTI eval cfg with rnn_modules / table capacity varied) -> (path, description, classes, frames, accumulate)"""
    desc, rnn, classes, frames, accumulate = WORKLOADS[name]
    if rnn is None and not accumulate:
        return CFG, desc, classes, frames, accumulate
    import hjson
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    if rnn is not None:
        cfg["model"]["rnn_modules"] = rnn
    if accumulate:
        cfg["lattice_gpu"]["hash_table_capacity"] = 200000
        cfg["loader_semantic_kitti"]["accumulate_clouds"] = True
    path = os.path.join("/tmp", "bench_%s_%d.cfg" % (name, os.getpid()))
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    return path, desc, classes, frames, accumulate


def config_of(args, desc, windows_np, scans_per_step):
    """the workload description both arms print verbatim"""
    return {"workload": desc, "points_per_scan": [[int(p.shape[0]) for p, _ in w] for w in windows_np], "scans_per_step": scans_per_step,
            "windows_cycled": len(windows_np), "window_seeds": "1000 + rank * windows + i"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_window(cfg, classes, window):
    """One window through the oracle's executor (oracle/window_oracle.py: the reference's model recipe in plain
    torch-CPU over the scalar C lattice oracle) with all host threads.  Returns (seconds, threads, log-softmax)."""
    import torch
    from oracle import window_oracle as WO
    torch.set_num_threads(os.cpu_count() or 1)
    runner = WO.OracleWindowRunner(cfg, classes)
    runner.materialise_parameters(window[:2])
    t0 = time.perf_counter()
    out = runner.infer_window(window)
    return time.perf_counter() - t0, torch.get_num_threads(), out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, desc, classes, frames, accumulate = workload_cfg(args.workload)
    windows = make_windows(args.windows, 1000, frames, accumulate)
    scans = FRAMES if accumulate else frames
    import torch
    from oracle import window_oracle as WO
    torch.set_num_threads(os.cpu_count() or 1)
    runner = WO.OracleWindowRunner(cfg, classes)
    runner.materialise_parameters(windows[0][:2])
    for i in range(args.warmup):
        runner.infer_window(windows[i % len(windows)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        runner.infer_window(windows[i % len(windows)])
    dt = time.perf_counter() - t0
    value = scans * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, desc, windows, scans),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d whole window(s) of the same synthetic workload (oracle/window_oracle.py)" % args.steps},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# HBM-side kernels of the metric
# ------------------------------------------------------------------------------------------------
def hbm_kernels(dev, window_np, pk):
    """The metric's second half: splat / slice / im2row (and distribute) HBM GB/s against the measured copy peak, on
    (a) ONE scan of the headline configuration and (b) the window's 4 scans accumulated into one cloud (~500k points,
    the accumulate_clouds shape of BASELINE config 5).  Every kernel is timed COLD: a CUDA graph alternates a 256 MB L2
    flush with the kernel 20 times, the time of the same graph with the flushes alone is subtracted.  Algorithmic bytes per
    SURVEY.md 8(d): compulsory traffic only."""
    import numpy as np
    import torch
    from temporal_latticenet_b200 import _lib
    from temporal_latticenet_b200.lattice import Lattice
    lib = _lib.load()
    p = _lib.ptr
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def graph_ms(fn, n=20):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            best = ms if best is None else min(best, ms)
        return best

    flush_ms = graph_ms(lambda: flush.zero_())

    def cold_ms(fn):
        def both():
            flush.zero_()
            fn()
        return max(graph_ms(both) - flush_ms, 1e-6)

    out = []

    def add(cloud, name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "cloud": cloud, "us": 1e3 * ms, "algorithmic_mb": nbytes / 1e6, "achieved": gbs, "unit": "GB/s",
                    "peak": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"], "shape": note, "l2": "flushed before every launch"})

    # context for the write-dominated kernels (slice writes N x C, im2row V x 9C): what a pure WRITE stream reaches on this device,
    # against the copy (read + write) figure the fractions are quoted on
    out.append({"kernel": "memset 256 MB (write-only reference, not a kernel of this library)", "cloud": "-", "us": 1e3 * flush_ms,
                "algorithmic_mb": (256 << 20) / 1e6, "achieved": (256 << 20) / (flush_ms * 1e-3) / 1e9, "unit": "GB/s", "peak": pk["hbm_gbs"],
                "frac": (256 << 20) / (flush_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "shape": "torch zero_ of the L2-flush buffer", "l2": "-"})
    clouds = [("one scan (headline config)", window_np[0][0], window_np[0][1]),
              ("4 scans accumulated", np.concatenate([f[0] for f in window_np], 0), np.concatenate([f[1] for f in window_np], 0))]
    for cloud, pos_np, val_np in clouds:
        pos, val = torch.from_numpy(np.ascontiguousarray(pos_np)).to(dev), torch.from_numpy(np.ascontiguousarray(val_np)).to(dev)
        N = pos.shape[0]
        ls = Lattice(100000, 0.6, device=dev)
        rows, idx, w = ls.distribute(pos, val, True)
        V = ls.nr_lattice_vertices()
        nbr = ls.neighbours()
        for C in (64, 192):
            feat = torch.randn(V, C, device=dev)
            rows_out = torch.empty(V, 9 * C, device=dev)
            add(cloud, "k_im2row", cold_ms(lambda: lib.ltn_im2row(p(feat), V, None, p(nbr), V, None, C, p(rows_out), _lib.stream())),
                V * (12 + 40 * C), "V=%d C=%d" % (V, C))
        for Cs in (32, 64, 192):
            vals = torch.randn(V, Cs, device=dev)
            sl = torch.empty(N, Cs, device=dev)
            add(cloud, "k_slice", cold_ms(lambda: lib.ltn_slice(p(vals), V, Cs, p(idx), p(w), N, p(sl), _lib.stream())),
                4 * N * 8 + V * 4 * Cs + N * 4 * Cs, "N=%d V=%d C=%d" % (N, V, Cs))
        for Cin in (1, 3):
            src = torch.randn(N, Cin, device=dev)
            acc = torch.zeros(V, Cin + 1, device=dev)
            add(cloud, "k_splat", cold_ms(lambda: lib.ltn_splat(p(src), N, Cin, p(idx), p(w), p(acc), V, _lib.stream())),
                N * (4 * Cin) + 4 * N * 8 + V * 4 * (Cin + 1), "N=%d V=%d C_in=%d (~%d rows per vertex)" % (N, V, Cin, 4 * N // max(V, 1)))
        ht = ls.hash_table

        def dist():
            lib.ltn_hash_clear(p(ht.slot_keys), p(ht.slot_ids), p(ht.slot_first), ht.nslots, p(ht.counters), _lib.stream())
            sx, sy, sz = ls.scale()
            lib.ltn_distribute(p(pos), p(val), N, None, 1, sx, sy, sz, p(ht.slot_keys), p(ht.slot_ids), p(ht.slot_first), ht.nslots,
                               p(ht.counters), p(ht.keys_tensor), ls.capacity, p(ls._row_slot), p(ls._block_sums), p(ls._vert_acc),
                               p(rows), p(idx), p(w), 1, _lib.stream())
        add(cloud, "ltn_distribute (table clear + hash build + rows + local mean)", cold_ms(dist), 128 * N, "N=%d -> V=%d" % (N, V))
    return out


# ------------------------------------------------------------------------------------------------
# training arm (BASELINE config 4)
# ------------------------------------------------------------------------------------------------
def run_train(args, rank, world, dev, devw, windows_np, lib, seeded_state):
    """BASELINE config 4: the same 4-frame gru-gru-aflow-gru window as a TRAINING step (BPTT through the 4
    frames, 0.5 Lovasz + 0.5 NLL, AdamW amsgrad), data-parallel over the ranks with one flattened
    gradient all-reduce per step."""
    import torch
    import torch.distributed as dist
    from temporal_latticenet_b200 import synthetic
    from temporal_latticenet_b200.train import WindowTrainer
    targets = [torch.from_numpy(synthetic.labels(w[-1][0].shape[0], 26, seed=i)).to(dev) for i, w in enumerate(windows_np)]
    tr = WindowTrainer(CFG, 26, dev)
    tr.materialise(devw[0], targets[0], seeded_state)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        tr.step(devw[i % len(devw)], targets[i % len(devw)])
    clocks = Clocks(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        clocks.start()
    barrier()
    l0 = lib.ltn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = tr.step(devw[i % len(devw)], targets[i % len(devw)])
    e1.record()
    barrier()
    launches = lib.ltn_launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        nparams = sum(p.numel() for p in tr.model.parameters())
        print(json.dumps({"metric": "4-frame scans/sec (training step)", "value": FRAMES * args.steps * world / (ms * 1e-3), "unit": UNIT,
                          "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "config4: same window, training step (BPTT, 0.5 Lovasz + 0.5 NLL, AdamW amsgrad), "
                                                 "data-parallel over windows, one flattened gradient all-reduce per step",
                                     "parameters": nparams, "allreduce_bytes_per_step": 4 * nparams,
                                     "allreduce_ms": getattr(tr, "allreduce_ms", None),
                                     "l2": "activations of a 4-frame window exceed L2"},
                          "gpu_launches": int(launches), "clocks": clk, "final_loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def stats_ms(xs):
    xs = sorted(xs)
    return {"min": xs[0], "median": xs[len(xs) // 2], "max": xs[-1], "n": len(xs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--windows", type=int, default=8, help="distinct synthetic windows cycled through (per rank; seeds 1000 + rank * windows + i): "
                    "enough of them that a rank's total work is an average over the window sizes, not the size of the two it happened to draw")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm", action="store_true", help="skip the splat / slice / im2row / distribute bandwidth table")
    ap.add_argument("--eager", action="store_true", help="eager per-op launches instead of CUDA-graph replay")
    ap.add_argument("--lanes", type=int, default=8, help="windows per lock-step group (one batched launch per layer serves them; at most 8). "
                    "When --steps is not a multiple of it, the remainder of every K-step pass runs as ONE smaller lock-step group")
    ap.add_argument("--groups", type=int, default=1, help="lock-step groups in flight per GPU (each with its own graphs and streams)")
    ap.add_argument("--streams", action="store_true", help="round-1 execution: one stream + one graph per window in flight "
                    "(MultiWindowRunner) instead of the lock-step group graph with batched tensor-core launches (LockstepRunner)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="the K-step timed loop is repeated until it has run this long")
    ap.add_argument("--profile-range", action="store_true",
                    help="after the timed regions, replay ONE more group between cudaProfilerStart/Stop: with `ncu --profile-from-start off` "
                         "the launch list then holds exactly the steady-state kernels of one group (profiles/README.md)")
    ap.add_argument("--driver", default="mirror", choices=["mirror", "reference"],
                    help="reference: ALSO time the reference's own seq_lattice/models.py over the shims (tools/reference_driver.py)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE config 3 (headline); train = config 4 (BPTT + AdamW + NCCL gradient all-reduce)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as G
    G.build()
    from temporal_latticenet_b200 import _lib
    from temporal_latticenet_b200.engine import GroupedLockstepRunner, LockstepRunner, MultiWindowRunner
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    lib = _lib.load()

    cfg, desc, classes, frames, accumulate = workload_cfg(args.workload)
    scans = FRAMES if accumulate else frames
    # every rank owns its OWN windows (sharded by window): per-rank spread of the window sizes is part of the number
    windows_np = make_windows(args.windows, 1000 + rank * args.windows, frames, accumulate)
    host = [[(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in w] for w in windows_np]
    devw = [[(p.to(dev), v.to(dev)) for p, v in w] for w in host]

    if args.mode == "train":
        return run_train(args, rank, world, dev, devw, windows_np, lib, seeded_state)
    lanes = max(1, args.lanes)
    # capacities are planned on every window of this rank and re-validated after every window (eager fallback when exceeded)
    if args.eager:
        runner = WindowRunner(cfg, classes, dev)
        runner.materialise_parameters(devw[0], seeded_state)
        multi, lanes = None, 1
    else:
        if args.streams:
            multi = MultiWindowRunner(cfg, classes, dev, lanes=lanes)
        elif args.groups > 1:
            multi = GroupedLockstepRunner(cfg, classes, dev, lanes=lanes, groups=args.groups)
            lanes = lanes * args.groups      # windows in flight per GPU
        else:
            lanes = min(lanes, 8, max(1, args.steps))
            multi = LockstepRunner(cfg, classes, dev, lanes=lanes)
        multi.prepare(devw[0], seeded_state, devw)
        runner = multi.lanes[0]
        if not multi.supported:
            multi, lanes = None, 1
    # K steps = full groups of `lanes` windows + ONE smaller lock-step group for the remainder (its own graphs), so that a pass
    # times exactly K windows whatever K is
    rem_multi, rem = None, args.steps % lanes
    if multi is not None and rem and isinstance(multi, LockstepRunner):
        rem_multi = LockstepRunner(cfg, classes, dev, lanes=rem)
        rem_multi.prepare(devw[0], seeded_state, devw)
        if not rem_multi.supported:
            rem_multi = None
    lockstep = isinstance(multi, (LockstepRunner, GroupedLockstepRunner))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nw = len(devw)

    def run_group(i0, n):
        """windows i0 .. i0+n-1 in flight together (one per lane); device-resident inputs"""
        if multi is None:
            for j in range(n):
                runner.infer_window_device(devw[(i0 + j) % nw])
        else:
            (rem_multi if (rem_multi is not None and n == rem) else multi).infer_windows_device([devw[(i0 + j) % nw] for j in range(n)])

    for i in range(0, max(args.warmup, 3), lanes):
        run_group(i, lanes)
    if rem_multi is not None:
        for _ in range(2):
            run_group(0, rem)
    torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs; EXACTLY `steps` windows per repeat, repeated until >= min_seconds ------
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    barrier()
    launches0 = lib.ltn_launch_count()
    evs, repeats, t_wall = [], 0, time.perf_counter()
    while True:
        for i in range(0, args.steps, lanes):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_group(i, min(lanes, args.steps - i))
            e1.record()
            evs.append((e0, e1, min(lanes, args.steps - i)))
        repeats += 1
        torch.cuda.synchronize()
        done = torch.tensor([1.0 if time.perf_counter() - t_wall >= args.min_seconds else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(done, op=dist.ReduceOp.MIN)   # every rank runs the same number of repeats
        if float(done.item()) > 0 or repeats >= 50:
            break
    barrier()
    launches = lib.ltn_launch_count() - launches0
    graph_mode = multi is not None
    capacity_ok = None
    if graph_mode:   # replayed graphs: the library's host-side counter saw the kernels once, at capture
        if lockstep:
            one = multi.groups[0] if hasattr(multi, "groups") else multi
            per = len(one.lanes)
            launches = sum((n // per) * one.kernels_per_group(frames) +
                           ((rem_multi.kernels_per_group(frames) if (rem_multi is not None and n % per == rem) else (n % per) * runner.kernels_per_window(frames))
                            if n % per else 0) for _, _, n in evs)
        else:
            launches = sum(n for _, _, n in evs) * runner.kernels_per_window(frames)
        capacity_ok = multi.counts_ok() and (rem_multi is None or rem_multi.counts_ok())
    group_ms = [a.elapsed_time(b) for a, b, _ in evs]
    ms_rank = sum(group_ms)
    t = torch.tensor([ms_rank], dtype=torch.float64, device=dev)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
    per_rank_ms = [float(x.item()) for x in per_rank]
    ms_max = max(per_rank_ms)
    timed_steps = args.steps * repeats

    # ---- timed region 2: end to end through the runner with pinned HOST buffers ----------------------
    def e2e_pass():
        labels = None
        if multi is None:
            for i in range(args.steps):
                labels = runner.infer_window(host[i % nw])
        else:   # one group of windows is always queued behind the one the host is waiting for
            pending = None
            for i in range(0, args.steps, lanes):
                n = min(lanes, args.steps - i)
                who = rem_multi if (rem_multi is not None and n == rem) else multi
                ticket = (who, who.submit([host[(i + j) % nw] for j in range(n)]))
                if pending is not None:
                    labels = pending[0].collect(pending[1])[-1]
                pending = ticket
            labels = pending[0].collect(pending[1])[-1]
        torch.cuda.synchronize()
        return labels
    e2e_pass()   # untimed: first use allocates the pinned label / counter buffers of this path
    barrier()
    t0 = time.perf_counter()
    for _ in range(repeats):
        labels = e2e_pass()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    fb = torch.zeros(2, dtype=torch.float64, device=dev)   # windows that left the graph path, summed over ALL ranks
    if multi is not None:
        all_lanes = list(multi.lanes) + (list(rem_multi.lanes) if rem_multi is not None else [])
        fb[0], fb[1] = sum(l.fallbacks for l in all_lanes), sum(l.range_fallbacks for l in all_lanes)
    if world > 1:
        dist.all_reduce(fb)
    if args.profile_range and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run_group(0, lanes)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    h2d = sum(sum(p.numel() * 4 + v.numel() * 4 for p, v in w) for w in host) // nw
    d2h = int(labels.numel() * 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # ---- roofline of the dominant kernel, measured INSIDE the replayed graphs of the timed configuration -------------
    roofline = roofline_of(multi if lockstep else None, runner, devw, frames, lanes, max(1, args.groups) if lockstep else 1, pk,
                           ms_max / timed_steps, lib, args)

    hbm = None
    if not args.no_hbm:
        try:
            hbm = hbm_kernels(dev, make_windows(1, 1000)[0], pk)
        except Exception as e:
            hbm = [{"failed": repr(e)}]

    # ---- CPU baseline (the oracle on the host cores) + label agreement of the timed window with it ---------------------
    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        try:
            sec, cores, want = cpu_window(cfg, classes, windows_np[0])
            cpu = {"value": scans / sec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "one whole window of the same workload (oracle/window_oracle.py), %.1f s" % sec}
            got = (multi.infer_windows_device([devw[j % nw] for j in range(lanes)])[0] if multi is not None
                   else runner.infer_window_device(devw[0])).float().cpu().numpy()
            want = want.numpy()
            fin = np.isfinite(want).all(1) & np.isfinite(got).all(1)
            agree = float((got[fin].argmax(1) == want[fin].argmax(1)).mean()) if fin.any() else None
            err = float(np.abs(got[fin] - want[fin]).max() / (np.abs(want[fin]).max() + 1e-30)) if fin.any() else None
            parity = {"window": "first timed window (seed 1000), last frame", "label_agreement_with_cpu_oracle": agree,
                      "log_softmax_max_err_over_absmax": err, "finite_rows": float(fin.mean()),
                      "nonfinite_mask_agreement": float((np.isfinite(want) == np.isfinite(got)).mean())}
            if agree is not None and agree < 0.999:
                raise AssertionError("the timed window's labels disagree with the CPU oracle: %r" % (parity,))
        except AssertionError:
            raise
        except Exception as e:  # the baseline leg must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}

    ref_driver = None
    if args.driver == "reference" and world == 1:
        try:
            r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "reference_driver.py"), "bench", "12", "3"],
                               capture_output=True, text=True, timeout=900)
            ref_driver = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else {"failed": r.stderr[-500:]}
        except Exception as e:
            ref_driver = {"failed": repr(e)}

    lvl = runner.static_lattice if getattr(runner, "caps", None) and graph_mode else runner.lattice
    v_counts = []
    while lvl is not None:
        v_counts.append(int(lvl.hash_table.count_tensor().cpu()))
        lvl = lvl._coarse
    total_scans = scans * timed_steps * world
    if lockstep:
        execution = ("%d lock-step group(s) of %d windows per GPU: ONE CUDA graph per frame kind covers a group, every tensor-core layer is one "
                     "persistent batched launch (k_conv_tc_batched), the other kernels run on per-window streams inside the graph"
                     % (max(1, args.groups), lanes // max(1, args.groups)))
        if rem_multi is not None:
            execution += "; the %d windows that remain of every %d-step pass run as one smaller lock-step group" % (rem, args.steps)
    elif graph_mode:
        execution = "CUDA-graph replay per frame kind, %d window(s) in flight per GPU (one stream + one graph each)" % lanes
    else:
        execution = "eager op-by-op launches"
    line = {"metric": METRIC, "value": total_scans / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / timed_steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, desc, windows_np, scans),
            "details": {"vertices_per_level_after_last_window": v_counts,
                        "precision": "fp32 results: every tensor-core product is the 3-pass hi/lo split (fp16 operands, exact power-of-two "
                                     "scaling, fp32 accumulation; range-flagged with a tf32 hi/lo re-run), 2e-5 * sum|a||w| of float64 in the tests",
                        "l2": "256 MB flush between groups; the per-window working set also exceeds L2",
                        "parallelism": "windows sharded over %d rank(s), each rank cycles its own %d windows, no data-path collective" % (world, nw),
                        "execution": execution, "static_capacities": getattr(runner, "caps", None), "capacities_respected": capacity_ok,
                        "timing": {"repeats_of_the_K_step_loop": repeats, "timed_steps_total": timed_steps, "timed_ms_total": ms_max,
                                   "ms_per_group_of_%d" % lanes: stats_ms([m for m, (_, _, n) in zip(group_ms, evs) if n == lanes] or group_ms),
                                   "per_rank_ms_total": per_rank_ms, "per_rank_ms_min_max": [min(per_rank_ms), max(per_rank_ms)]}},
            "e2e": {"value": total_scans / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "hbm_kernels": hbm, "cpu_baseline": cpu, "parity": parity}
    if ref_driver is not None:
        line["reference_driver"] = ref_driver
    if graph_mode:   # windows that left the graph path: static capacities exceeded / fp16 operand range flag raised
        line["fallbacks"] = {"eager_reruns": int(fb[0].item()), "fp16_range": int(fb[1].item()), "scope": "all ranks"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline_of(lockstep_runner, runner, devw, frames, lanes, groups, pk, ms_per_step, lib, args):
    """Tensor roofline of the dominant kernel.  Lock-step mode: every batched launch stamps %globaltimer at the entry and
    exit of each CTA into a trace buffer (csrc/ltn_conv_batched.cu), so the durations are those of the launches INSIDE
    the replayed graphs of the timed configuration -- with the other windows' small kernels running beside them, warm
    weights, real data.  duration of a launch = last CTA exit - first CTA entry."""
    note3 = ("the parity mode issues 3 MMAs per product (fp16 operands at the bf16 rate), so 1/3 of the bf16 peak is this kernel's ceiling")
    traffic, traffic_note = None, None
    for name in ("r2_conv_batched_192_ncu_full.json", "r1_conv_tc_f16_192_ncu_full.json"):
        tpath = os.path.join(REPO, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["traffic_bytes"]
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (%s) from ncu --set full (profiles/%s): %.2f MB "
                            "against %.2f MB algorithmic for that launch" % (tj["launch"], name, tj["traffic_bytes"] / 1e6, tj["algorithmic_bytes"] / 1e6))
            break
    if lockstep_runner is None or not hasattr(lockstep_runner, "trace_group"):
        return {"kernel": "k_conv_tc", "bound": "tensor", "achieved": None, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": None,
                "traffic": traffic, "note": "per-launch timing is only available in the lock-step execution mode"}
    nw = len(devw)
    rec = lockstep_runner.trace_group([devw[j % nw] for j in range(lanes)])
    if not rec:
        return None
    total_us = sum(r["us"] for r in rec)
    total_fl = sum(r["flop"] for r in rec)
    by_shape = {}
    for r in rec:
        k = "C%d S%d F%d rows~%dk" % (r["C"], r["S"], r["F"], round(r["rows"] / 1000.0))
        e = by_shape.setdefault(k, {"launches": 0, "us": 0.0, "flop": 0.0})
        e["launches"] += 1
        e["us"] += r["us"]
        e["flop"] += r["flop"]
    shapes = {k: {"launches": v["launches"], "us_total": v["us"], "tflops": v["flop"] / (v["us"] * 1e-6) / 1e12 if v["us"] > 0 else None,
                  "frac_of_bf16_sustained": v["flop"] / (v["us"] * 1e-6) / 1e12 / pk["bf16_tflops_sustained"] if v["us"] > 0 else None}
              for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1]["us"])}
    achieved = total_fl / (total_us * 1e-6) / 1e12
    ceiling = pk["bf16_tflops_sustained"] / 3.0
    per = lanes // groups    # windows whose launches `rec` holds (one lock-step group)
    step_tf = (total_fl / per) / (ms_per_step * 1e-3) / 1e12
    return {"kernel": "k_conv_tc_batched (fused gather + GroupNorm/ReLU + tcgen05 GEMM, fp32-parity 3-pass fp16 hi/lo split; one persistent "
                      "launch per layer for the %d windows of a lock-step group)" % per,
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
            "traffic": traffic, "traffic_note": traffic_note, "peak_source": pk["source"] + " (bf16 sustained)",
            "frac_of_3pass_ceiling": achieved / ceiling, "ceiling_3pass": ceiling,
            "note": "achieved = algorithmic 2*V*S*C*F flop of every batched launch of one group of %d windows / the sum of their in-graph durations "
                    "(globaltimer stamps of the first CTA entry and last CTA exit); %s" % (per, note3),
            "launches_per_group": len(rec), "avg_launch_us": total_us / len(rec),
            "kernel_ms_per_group": total_us / 1e3, "share_of_step": (total_us / 1e3 / per) / ms_per_step,
            "share_note": "kernel time per window / ms_per_step; with several groups in flight the groups' launches overlap, so this can exceed 1/groups but never 1",
            "whole_step": {"gflop_per_window": total_fl / per / 1e9, "tflops": step_tf, "frac": step_tf / pk["bf16_tflops_sustained"],
                           "note": "algorithmic conv flop of a window / ms_per_step: what the whole step achieves, everything included"},
            "tensor_flops_issued_tflops": 3 * achieved, "per_shape": shapes}


if __name__ == "__main__":
    main()
