"""Shared test helpers: seeded state-dicts by parameter name, small synthetic windows, canonical
vertex order."""
import os
import sys
import zlib

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
CFG = os.path.join(REPO, "configs", "lnn_eval_semantic_kitti.cfg")


def seeded_tensor(name, shape):
    """Deterministic value for a parameter, a function of its NAME and SHAPE only (the pretrained
    checkpoint is missing from the reference mount -- .MISSING_LARGE_BLOBS:1 -- so oracle and CUDA
    paths both load this instead)."""
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    shape = tuple(shape)
    if len(shape) == 0:
        return torch.tensor(0.1)  # AFlow alpha / beta (lattice_modules.py:252-253)
    x = torch.randn(shape, generator=g)
    if name.endswith("gn.weight") or name.endswith(".gamma"):
        return 1.0 + 0.1 * x
    if len(shape) == 1:
        return 0.1 * x
    return x * float(np.sqrt(2.0 / max(shape)))


def seeded_state(shapes):
    return {k: seeded_tensor(k, v) for k, v in sorted(shapes.items())}


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
