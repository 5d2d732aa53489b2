"""Deterministic stand-in for the pretrained checkpoint (missing from the reference mount,
.MISSING_LARGE_BLOBS:1): every parameter is a function of its NAME and SHAPE only, so the CPU run of
the reference's model code (tests/golden/make_golden.py), the oracle and the CUDA path all load
identical weights through `load_state_dict`."""
import zlib

import numpy as np
import torch


def seeded_tensor(name, shape):
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
