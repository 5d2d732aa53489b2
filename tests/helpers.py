"""Shared test helpers: seeded state-dicts by parameter name, small synthetic windows, canonical
vertex order."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
CFG = os.path.join(REPO, "configs", "lnn_eval_semantic_kitti.cfg")


from temporal_latticenet_b200.seeding import seeded_state, seeded_tensor  # noqa: E402,F401


def small_window(seed=1, frames=4, radius=9.0, max_points=6000):
    """Crop of a synthetic window around the sensor: dense enough that most vertices pass the
    min-4-points mask (lattice_modules.py:528), small enough for the CPU oracle."""
    from temporal_latticenet_b200 import synthetic
    out = []
    for f, (p, v) in enumerate(synthetic.window(seed, frames=frames)):
        keep = np.linalg.norm(p[:, [0, 2]], axis=1) < radius
        p, v = p[keep], v[keep]
        if p.shape[0] > max_points:
            sel = np.sort(np.random.default_rng(seed * 17 + f).choice(p.shape[0], max_points, replace=False))
            p, v = p[sel], v[sel]
        out.append((np.ascontiguousarray(p), np.ascontiguousarray(v)))
    return out


def canonical_order(keys):
    """permutation that sorts [V,3] integer keys lexicographically"""
    keys = np.asarray(keys)
    return np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
