"""The HBM-side kernels the metric names (im2row, slice, splat, distribute) on the accumulated 4-scan cloud: the same
measurement bench.py reports under `hbm_kernels`, stand-alone.  python tools/bench_hbm.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as G
G.build()
dev = torch.device("cuda:0")
win = bench.make_windows(1, 1000)[0]
for h in bench.hbm_kernels(dev, win, bench.peaks()):
    print("%-40s %7.1f us  %7.1f MB  %7.1f GB/s  frac %.3f  %s" % (h["kernel"][:40], h["us"], h["algorithmic_mb"], h["achieved"], h["frac"], h["shape"]))
