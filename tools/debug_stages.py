#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA path with the CPU oracle on one window: forward hooks on the sub-modules the two
models share by name (state-dict compatible), per frame.  Test infrastructure (imports oracle/).
   python tools/debug_stages.py [frames] [seed] [crop_radius or 0] [rnn, comma separated]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tests.helpers import CFG  # noqa: E402


def hook_all(model, store):
    names = ["point_net_seq"]
    for n, m in model.named_modules():
        parts = n.split(".")
        if (parts[0] in ("resnet_blocks_per_down_lvl_list", "resnet_blocks_per_up_lvl_list") and len(parts) == 3) or \
           (parts[0] in ("coarsens_list", "finefy_list", "resnet_blocks_bottleneck", "recurrent_fusion_modules") and len(parts) == 2) or \
           n in ("point_net_seq", "point_net_seq.fusion_module", "point_net_seq.last_conv", "slice_fast_cuda") or n.endswith(".AFLOW"):
            def mk(name):
                def hook(mod, inp, out):
                    t = out[0] if isinstance(out, tuple) else out
                    store.append((name, t.detach().cpu().numpy().copy()))
                    if name.endswith(".AFLOW"):
                        store.append((name + ".weights", out[1].detach().cpu().numpy().copy()))
                        store.append((name + ".nbr", out[2].detach().cpu().numpy().astype(np.float64)))
                        store.append((name + ".in_lv", inp[0].detach().cpu().numpy().copy()))
                        store.append((name + ".in_h", inp[1].detach().cpu().numpy()[: min(1571, inp[1].shape[0])].copy()))
                return hook
            m.register_forward_hook(mk(n))


def main():
    frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    radius = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    rnn = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    from temporal_latticenet_b200 import ops, synthetic
    from temporal_latticenet_b200.runner import WindowRunner
    from temporal_latticenet_b200.seeding import seeded_state
    from oracle import window_oracle as WO
    cfg = CFG
    if rnn:
        import hjson
        with open(CFG) as f:
            c = hjson.loads(f.read())
        c["model"]["rnn_modules"] = rnn
        cfg = "/tmp/debug_stages.cfg"
        with open(cfg, "w") as f:
            f.write(hjson.dumps(c))
    window = synthetic.window(seed, frames=frames_n, scope=3)
    if radius > 0:
        window = [(np.ascontiguousarray(p[np.linalg.norm(p[:, [0, 2]], axis=1) < radius]),
                   np.ascontiguousarray(v[np.linalg.norm(p[:, [0, 2]], axis=1) < radius])) for p, v in window]
    dev = torch.device("cuda:0")
    fd = [(torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)) for p, v in window]
    operands = os.environ.get("DBG_OPERANDS", "f16")
    run = WindowRunner(cfg, 26, dev, operands=operands).materialise_parameters(fd, seeded_state)
    orc = WO.OracleWindowRunner(cfg, 26).materialise_parameters(window[:2])
    got, want = [], []
    hook_all(run.model, got)
    hook_all(orc.model, want)
    run.infer_window_device(fd)
    orc.infer_window(window)
    print("hooks: gpu %d oracle %d" % (len(got), len(want)))
    for (n1, a), (n2, b) in zip(got, want):
        assert n1 == n2, (n1, n2)
        if a.shape != b.shape:
            print("%-50s SHAPE %s vs %s" % (n1, a.shape, b.shape))
            continue
        fin = np.isfinite(b)
        scale = float(np.abs(b[fin]).max()) + 1e-30 if fin.any() else 1.0
        err = np.abs(a.astype(np.float64) - b)[fin] / scale
        rows_bad = int((np.abs(a.astype(np.float64) - b).max(1) > 1e-4 * scale).sum()) if a.ndim == 2 else -1
        if n1.endswith(".weights"):
            bad = np.argwhere(np.abs(a - b) > 1e-5)
            print("   weights differing > 1e-5: %d entries; first:" % len(bad), [(int(r), int(c), float(a[r, c]), float(b[r, c])) for r, c in bad[:12]])
        print("%-50s %-14s absmax %.3e  med %.1e  p99.9 %.1e  max %.1e  rows>1e-4: %d  nan: gpu %d oracle %d" % (
            n1, a.shape, scale, float(np.median(err)) if err.size else 0, float(np.quantile(err, 0.999)) if err.size else 0,
            float(err.max()) if err.size else 0, rows_bad, int((~np.isfinite(a)).sum()), int((~fin).sum())))


if __name__ == "__main__":
    main()
