"""`latticenet`-compatible host objects over the sm_100a kernels: ModelParams, HashTable, Lattice.

Mirrors what the reference imports from the absent C++ extension
(/root/reference/seq_lattice/lattice_modules.py:7-8, train_ln.py:16,80,106): same names, same
argument meaning; glog CHECK aborts become RuntimeError.  Device memory is plain torch tensors,
kernels are reached through the C ABI in include/latticenet_b200.h on the current stream.
"""
import math

import hjson
import torch

from . import _lib

POS_DIM = 3
FILTER_EXTENT = 9
# convention U1: Adams' (d+1)*sqrt(2/3); see DESIGN.md
INV_STD_DEV_ADAMS = (POS_DIM + 1) * math.sqrt(2.0 / 3.0)


class ModelParams:
    """Getter struct over the cfg `model` section (train_ln.py:80; models.py:29-37,63-64,488,503)."""

    def __init__(self, d):
        self._d = dict(d)

    @staticmethod
    def create(cfg_path):
        with open(cfg_path, "r") as f:
            return ModelParams(hjson.loads(f.read())["model"])

    def positions_mode(self): return str(self._d["positions_mode"])
    def values_mode(self): return str(self._d["values_mode"])
    def pointnet_layers(self): return [int(x) for x in self._d["pointnet_layers"]]
    def pointnet_start_nr_channels(self): return int(self._d["pointnet_start_nr_channels"])
    def nr_downsamples(self): return int(self._d["nr_downsamples"])
    def nr_blocks_down_stage(self): return [int(x) for x in self._d["nr_blocks_down_stage"]]
    def nr_blocks_bottleneck(self): return int(self._d["nr_blocks_bottleneck"])
    def nr_blocks_up_stage(self): return [int(x) for x in self._d["nr_blocks_up_stage"]]
    def nr_levels_down_with_normal_resnet(self): return int(self._d["nr_levels_down_with_normal_resnet"])
    def nr_levels_up_with_normal_resnet(self): return int(self._d["nr_levels_up_with_normal_resnet"])
    def compression_factor(self): return float(self._d["compression_factor"])
    def dropout_last_layer(self): return float(self._d["dropout_last_layer"])
    def experiment(self): return str(self._d["experiment"])


def scale_factors(sigma, inv_std_dev=INV_STD_DEV_ADAMS):
    """Three fp32 scale factors, computed in double and rounded once."""
    import numpy as np
    return [float(np.float32(inv_std_dev / (float(sigma) * math.sqrt((i + 1) * (i + 2))))) for i in range(POS_DIM)]


class HashTable:
    """Device hash table: packed 63-bit keys claimed with one CAS, vertex ids numbered in order of
    first appearance (see csrc/ltn_lattice.cu)."""

    def __init__(self, capacity, device=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self._capacity = int(capacity)
        # 4 slots per vertex: probes stay short, and a frame may touch up to 4x capacity distinct keys
        # before the overflow rule (ids beyond capacity -> -1, in order of first appearance) degrades
        n = 8
        while n < 4 * self._capacity:
            n <<= 1
        self.nslots = n
        dev = self.device
        self.slot_keys = torch.empty(n, dtype=torch.int64, device=dev)
        self.slot_ids = torch.empty(n, dtype=torch.int32, device=dev)
        self.slot_first = torch.empty(n, dtype=torch.int32, device=dev)
        self.keys_tensor = torch.zeros(self._capacity, 4, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(8, dtype=torch.int32, device=dev)
        self._host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._host_valid = False
        self.static_rows = None
        self.clear()

    def capacity(self):
        return self._capacity

    def clear(self):
        _lib.check(self.lib.ltn_hash_clear(_lib.ptr(self.slot_keys), _lib.ptr(self.slot_ids), _lib.ptr(self.slot_first),
                                           self.nslots, _lib.ptr(self.counters), _lib.stream()), "ltn_hash_clear")
        self._host_valid = False

    def invalidate(self):
        self._host_valid = False

    def _sync_counters(self):
        if not self._host_valid:
            self._host.copy_(self.counters, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self._host_valid = True
            if int(self._host[3]) != 0:
                raise RuntimeError("lattice keys left the packable +-2^20 range (positions/sigma too large)")
        return self._host

    def nr_filled(self):
        """Device -> host read of the vertex counter (one stream sync, cached until the next insert).
        In static-capacity mode (engine.py) the answer is the level's fixed capacity and nothing syncs."""
        if self.static_rows is not None:
            return self.static_rows
        return int(self._sync_counters()[0])

    def count_tensor(self):
        """device int32 [1]: the live vertex count"""
        return self.counters[0:1]

    def nr_overflowed(self):
        return int(self._sync_counters()[2])

    def keys(self):
        """[V,3] int32 keys in vertex-id order"""
        return self.keys_tensor[: int(self._sync_counters()[0]), :3]


class Lattice:
    """Stateful lattice handle (`ls` in the reference).  One object per resolution level; the coarser
    level hangs off `_coarse` and persists for the lifetime of the root so coarse vertex ids are
    append-only across the frames of a window (convention U3 / quirk Q2)."""

    def __init__(self, capacity, sigma, level=0, inv_std_dev=INV_STD_DEV_ADAMS, device=None):
        self.capacity = int(capacity)
        self.sigma = float(sigma)
        self.level = level
        self.inv_std_dev = inv_std_dev
        self.hash_table = HashTable(capacity, device)
        self.device = self.hash_table.device
        self.lib = self.hash_table.lib
        self._values = None
        self._positions = None
        self._frame = 0
        self._coarse = None
        self._coarse_frame = -1
        self._nbr_cache = {}
        self._scratch_n = 0
        self._row_slot = None
        self._block_sums = None
        self._vert_acc = None
        self.vertex_counts = None
        self.static_caps = None   # engine.py: per-level vertex capacities [V0cap, V1cap, ...]

    # ---- reference-visible API ---------------------------------------------------------------
    @staticmethod
    def create(cfg_path, name="lattice"):
        with open(cfg_path, "r") as f:
            cfg = hjson.loads(f.read())["lattice_gpu"]
        if int(cfg["nr_sigmas"]) != 1:
            raise RuntimeError("only one sigma group is supported (the reference cfgs use nr_sigmas: 1)")
        val, extent = str(cfg["sigma_0"]).split()
        if int(extent) != POS_DIM:
            raise RuntimeError("sigma_0 must cover pos_dim = 3 dimensions")
        return Lattice(int(cfg["hash_table_capacity"]), float(val))

    def set_values(self, v):
        """O(1) rebinding (callers do this all over: lattice_modules.py:38...574)."""
        self._values = v

    def values(self): return self._values
    def val_dim(self): return int(self._values.shape[1])
    def pos_dim(self): return POS_DIM
    def positions(self): return self._positions
    def nr_lattice_vertices(self): return self.hash_table.nr_filled()

    def get_filter_extent(self, neighbourhood_size):
        if neighbourhood_size != 1:
            raise RuntimeError("only the 1-hop neighbourhood is supported (lattice_modules.py:299)")
        return FILTER_EXTENT

    # ---- structure ops -----------------------------------------------------------------------
    def scale(self):
        return scale_factors(self.sigma, self.inv_std_dev)

    def _bump(self):
        self._frame += 1
        self._nbr_cache = {}
        self.hash_table.invalidate()

    def _scratch(self, n_points):
        if self._row_slot is None or self._scratch_n < n_points:
            self._scratch_n = int(n_points)
            self._row_slot = torch.empty(4 * self._scratch_n, dtype=torch.int32, device=self.device)
            self._block_sums = torch.empty((4 * self._scratch_n + 1023) // 1024 + 1, dtype=torch.int32, device=self.device)
        if self._vert_acc is None:
            self._vert_acc = torch.empty(self.capacity, 4, dtype=torch.float64, device=self.device)

    def _check_points(self, positions):
        if positions.dim() != 2 or positions.shape[1] != POS_DIM:
            raise RuntimeError("positions must be [N,3]")
        if positions.dtype != torch.float32 or not positions.is_cuda:
            raise RuntimeError("positions must be a float32 CUDA tensor")
        return positions.contiguous()

    def distribute(self, positions, values, reset_hashmap=True, subtract_mean=True):
        """(rows [4N, 3+vd+1], idx [4N] i32, w [4N]) -- models.py:297-298"""
        positions = self._check_points(positions)
        values = values.contiguous().float()
        if values.dim() != 2 or values.shape[0] != positions.shape[0]:
            raise RuntimeError("values must be [N, val_dim]")
        if reset_hashmap:
            self.reset()
        self._bump()
        self._positions = positions
        n, vd = positions.shape[0], values.shape[1]
        self._scratch(n)
        rows = torch.empty(4 * n, POS_DIM + vd + 1, dtype=torch.float32, device=self.device)
        idx = torch.empty(4 * n, dtype=torch.int32, device=self.device)
        w = torch.empty(4 * n, dtype=torch.float32, device=self.device)
        sx, sy, sz = self.scale()
        ht = self.hash_table
        p = _lib.ptr
        _lib.check(self.lib.ltn_distribute(p(positions), p(values), n, _lib.rows_dev(n), vd, sx, sy, sz, p(ht.slot_keys), p(ht.slot_ids),
                                           p(ht.slot_first), ht.nslots, p(ht.counters), p(ht.keys_tensor), self.capacity,
                                           p(self._row_slot), p(self._block_sums), p(self._vert_acc), p(rows), p(idx),
                                           p(w), 1 if subtract_mean else 0, _lib.stream()), "ltn_distribute")
        return rows, idx, w

    def rows_per_vertex(self, nr_vertices):
        """float [V]: number of distributed rows per vertex (ids < 0 folded onto vertex 0), from the
        accumulator the last distribute left behind (lattice_modules.py:519-521)."""
        out = torch.empty(nr_vertices, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ltn_vertex_counts(_lib.ptr(self._vert_acc), nr_vertices, _lib.ptr(out), _lib.stream()),
                   "ltn_vertex_counts")
        return out

    def just_create_verts(self, positions, reset_hashmap=True):
        positions = self._check_points(positions)
        if reset_hashmap:
            self.reset()
        self._bump()
        self._positions = positions
        n = positions.shape[0]
        self._scratch(n)
        sx, sy, sz = self.scale()
        ht = self.hash_table
        p = _lib.ptr
        _lib.check(self.lib.ltn_insert_points(p(positions), n, _lib.rows_dev(n), sx, sy, sz, p(ht.slot_keys), p(ht.slot_ids),
                                              p(ht.slot_first), ht.nslots, p(ht.counters), p(ht.keys_tensor),
                                              self.capacity, p(self._row_slot), None, p(self._block_sums),
                                              _lib.stream()), "ltn_insert_points")

    def reset(self):
        """begin a new sequence: empty this level's table; the coarser levels are dropped (eager mode) or,
        in static-capacity mode, kept as objects and emptied too so captured graphs keep their addresses"""
        self.hash_table.clear()
        if self.static_caps is None:
            self._coarse = None
        elif self._coarse is not None:
            self._coarse.reset()
        self._coarse_frame = -1

    def set_static(self, caps):
        """static-capacity mode: caps[level] = fixed row count of every per-vertex tensor of that level"""
        self.static_caps = None if caps is None else list(caps)
        self.hash_table.static_rows = None if caps is None else int(caps[self.level])
        if self._coarse is not None:
            self._coarse.set_static(caps)

    def coarse_level(self):
        """the (persistent) next-coarser lattice object, created empty on first use"""
        if self._coarse is None:
            self._coarse = Lattice(self.capacity, self.sigma * 2.0, self.level + 1, self.inv_std_dev, self.device)
            self._coarse.set_static(self.static_caps)
        return self._coarse

    def create_coarse_verts(self):
        """Coarser lattice (sigma x2) holding every coarse vertex touched by any frame so far."""
        self.coarse_level()
        c = self._coarse
        if self._coarse_frame != self._frame:
            c.just_create_verts(self._positions, reset_hashmap=False)
            self._coarse_frame = self._frame
        return c

    def neighbours(self, other=None, mode=0, dilation=1):
        """[V,9] int32 ids in `other` (default self); -1 absent; slot 8 = centre.  Built once per
        lattice state and shared by every convolution on the level."""
        other = other or self
        key = (id(other), other._frame, mode, dilation)
        t = self._nbr_cache.get(key)
        if t is None:
            v = self.nr_lattice_vertices()
            t = torch.empty(v, FILTER_EXTENT, dtype=torch.int32, device=self.device)
            ht = other.hash_table
            p = _lib.ptr
            _lib.check(self.lib.ltn_neighbours(p(self.hash_table.keys_tensor), v, _lib.rows_dev(v), p(ht.slot_keys), p(ht.slot_ids),
                                               ht.nslots, mode, dilation, 1 if other is self else 0, p(t),
                                               _lib.stream()), "ltn_neighbours")
            self._nbr_cache[key] = t
        return t
