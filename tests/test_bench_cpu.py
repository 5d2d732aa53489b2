"""Host-side contract of bench.py that the driver relies on (no GPU): both arms describe the workload with the SAME `config`
object, ranks draw distinct windows, the JSON of the reference arm carries the keys the tier asks for."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def test_both_arms_print_the_same_config_and_ranks_draw_distinct_windows():
    args = argparse.Namespace(windows=2)
    cfg, desc, classes, frames, accumulate = bench.workload_cfg("config3")
    w0 = bench.make_windows(2, 1000, frames, accumulate)          # rank 0 of our arm == the reference arm's windows
    w0_again = bench.make_windows(2, 1000, frames, accumulate)
    w1 = bench.make_windows(2, 1000 + 1 * 2, frames, accumulate)  # rank 1
    a = bench.config_of(args, desc, w0, frames)
    b = bench.config_of(args, desc, w0_again, frames)
    assert a == b and set(a) == {"workload", "points_per_scan", "scans_per_step", "windows_cycled", "window_seeds"}
    assert a["workload"].startswith("config3") and a["scans_per_step"] == 4 and a["windows_cycled"] == 2
    assert all(np.array_equal(p, q) for (p, _), (q, _) in zip(w0[0], w0_again[0]))        # seeded: reproducible
    assert not np.array_equal(w0[0][0][0][:100], w1[0][0][0][:100])                        # another rank, other windows
    assert all(110000 < n < 140000 for w in a["points_per_scan"] for n in w)              # "about 120k points" per scan


def test_accumulated_workload_is_one_cloud_of_four_scans():
    cfg, desc, classes, frames, accumulate = bench.workload_cfg("config5-accumulated")
    assert accumulate and frames == 1 and classes == 26
    import hjson
    with open(cfg) as f:
        c = hjson.loads(f.read())
    assert c["model"]["rnn_modules"] == ["aflow"] * 4 and c["loader_semantic_kitti"]["accumulate_clouds"] is True
    w = bench.make_windows(1, 1000, frames, accumulate)[0]
    assert len(w) == 1 and 450000 < w[0][0].shape[0] < 560000   # ~480k points per lattice (BASELINE config 5)


def test_bench_refuses_to_run_our_arm_without_a_gpu():
    """no CPU fallback: the product arm must fail loudly where there is no CUDA device"""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
