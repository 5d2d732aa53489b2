"""ctypes front-end of the scalar C oracle (oracle/lattice_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
PARITY UNPINNED (see the header of lattice_oracle.c): the reference ships no golden vectors for
the `latticenet` boundary; the oracle restates the published permutohedral algorithm and the
call-site contracts of /root/reference/seq_lattice/{models,lattice_modules}.py.
"""
import ctypes
import math
import os

import numpy as np

from . import build as _build

_lib = None

D = 3
D1 = 4
FEXT = 9

# convention U1: Adams' (d+1)*sqrt(2/3); the upstream project is recalled to
# hard-code 1.0 but the reference's own evidence (cfg comment "around 10k with sigma of 1",
# seq_config/lnn_train_semantic_kitti.cfg:71) only fits Adams' factor.  Exposed as a parameter.
INV_STD_DEV_ADAMS = (D + 1) * math.sqrt(2.0 / 3.0)


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.orc_table_create.restype = vp
        L.orc_table_create.argtypes = [ci]
        L.orc_table_free.argtypes = [vp]
        L.orc_table_clear.argtypes = [vp]
        L.orc_table_size.restype = ci
        L.orc_table_size.argtypes = [vp]
        L.orc_table_capacity.restype = ci
        L.orc_table_capacity.argtypes = [vp]
        L.orc_table_keys.restype = ctypes.POINTER(ctypes.c_int)
        L.orc_table_keys.argtypes = [vp]
        L.orc_table_find.restype = ci
        L.orc_table_find.argtypes = [vp, vp]
        L.orc_table_insert.restype = ci
        L.orc_table_insert.argtypes = [vp, vp]
        L.orc_simplex.argtypes = [vp, vp, vp, vp]
        L.orc_distribute.argtypes = [vp, vp, vp, ci, ci, vp, vp, vp, vp]
        L.orc_insert_points.argtypes = [vp, vp, ci, vp]
        L.orc_local_mean_sub.argtypes = [vp, vp, ci, ci, ci]
        L.orc_neighbours.argtypes = [vp, ci, vp, ci, ci, vp]
        L.orc_im2row.argtypes = [vp, ci, vp, ci, ci, vp]
        L.orc_splat.argtypes = [vp, ci, ci, vp, vp, vp, ci]
        L.orc_slice.argtypes = [vp, ci, ci, vp, vp, ci, vp]
        L.orc_gather.argtypes = [vp, ci, ci, vp, vp, ci, vp]
        L.orc_slice_classify.argtypes = [vp, ci, ci, vp, vp, vp, ci, vp, vp, ci, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def scale_factors(sigma, inv_std_dev=INV_STD_DEV_ADAMS):
    """scale[i] = inv_std_dev / (sigma * sqrt((i+1)(i+2))), computed in double, rounded once to fp32
    (SURVEY.md appendix B.1).  The product computes the identical three floats on the host."""
    return np.array([inv_std_dev / (float(sigma) * math.sqrt((i + 1) * (i + 2))) for i in range(D)],
                    dtype=np.float32)


class OracleTable:
    """Open-addressing hash of lattice keys -> dense vertex ids (insertion order)."""

    def __init__(self, capacity):
        self._h = lib().orc_table_create(int(capacity))
        self.capacity = int(capacity)

    def __del__(self):
        try:
            if self._h:
                lib().orc_table_free(self._h)
                self._h = None
        except Exception:
            pass

    def clear(self):
        lib().orc_table_clear(self._h)

    def size(self):
        return lib().orc_table_size(self._h)

    def keys(self):
        n = self.size()
        ptr = lib().orc_table_keys(self._h)
        return np.ctypeslib.as_array(ptr, shape=(self.capacity, D))[:n].copy()

    def find(self, key):
        k = _i32(key)
        return lib().orc_table_find(self._h, _p(k))

    def insert(self, key):
        k = _i32(key)
        return lib().orc_table_insert(self._h, _p(k))

    def distribute(self, pos, val, scale):
        pos, val, scale = _f32(pos), _f32(val), _f32(scale)
        n, vd = pos.shape[0], val.shape[1]
        rows = np.zeros((n * D1, D + vd + 1), np.float32)
        idx = np.full((n * D1,), -1, np.int32)
        w = np.zeros((n * D1,), np.float32)
        lib().orc_distribute(self._h, _p(pos), _p(val), n, vd, _p(scale), _p(rows), _p(idx), _p(w))
        return rows, idx, w

    def insert_points(self, pos, scale):
        pos, scale = _f32(pos), _f32(scale)
        lib().orc_insert_points(self._h, _p(pos), pos.shape[0], _p(scale))

    def neighbours(self, nbr_table=None, mode=0, dilation=1, nr_query=None):
        nbr_table = nbr_table or self
        vq = self.size() if nr_query is None else int(nr_query)
        out = np.full((vq, FEXT), -1, np.int32)
        lib().orc_neighbours(self._h, vq, nbr_table._h, int(mode), int(dilation), _p(out))
        return out


def simplex(p, scale):
    p, scale = _f32(p), _f32(scale)
    keys = np.zeros((D1, D), np.int32)
    bary = np.zeros((D1,), np.float32)
    lib().orc_simplex(_p(p), _p(scale), _p(keys), _p(bary))
    return keys, bary


def local_mean_sub(rows, idx, nr_vertices):
    rows = _f32(rows).copy()
    idx = _i32(idx)
    lib().orc_local_mean_sub(_p(rows), _p(idx), rows.shape[0], rows.shape[1], int(nr_vertices))
    return rows


def im2row(nbr, vals):
    nbr, vals = _i32(nbr), _f32(vals)
    out = np.zeros((nbr.shape[0], FEXT * vals.shape[1]), np.float32)
    lib().orc_im2row(_p(nbr), nbr.shape[0], _p(vals), vals.shape[0], vals.shape[1], _p(out))
    return out


def splat(val, idx, w, nr_vertices):
    val, idx, w = _f32(val), _i32(idx), _f32(w)
    out = np.zeros((nr_vertices, val.shape[1] + 1), np.float32)
    lib().orc_splat(_p(val), val.shape[0], val.shape[1], _p(idx), _p(w), _p(out), int(nr_vertices))
    return out


def slice_(vals, idx, w):
    vals, idx, w = _f32(vals), _i32(idx), _f32(w)
    n = idx.shape[0] // D1
    out = np.zeros((n, vals.shape[1]), np.float32)
    lib().orc_slice(_p(vals), vals.shape[0], vals.shape[1], _p(idx), _p(w), n, _p(out))
    return out


def gather(vals, idx, w):
    vals, idx, w = _f32(vals), _i32(idx), _f32(w)
    n = idx.shape[0] // D1
    out = np.zeros((n, D1 * (vals.shape[1] + 1)), np.float32)
    lib().orc_gather(_p(vals), vals.shape[0], vals.shape[1], _p(idx), _p(w), n, _p(out))
    return out


def slice_classify(vals, idx, w, dw, W, b):
    vals, idx, w, dw, W, b = _f32(vals), _i32(idx), _f32(w), _f32(dw), _f32(W), _f32(b)
    n = idx.shape[0] // D1
    out = np.zeros((n, W.shape[0]), np.float32)
    lib().orc_slice_classify(_p(vals), vals.shape[0], vals.shape[1], _p(idx), _p(w), _p(dw), n,
                             _p(W), _p(b), W.shape[0], _p(out))
    return out
