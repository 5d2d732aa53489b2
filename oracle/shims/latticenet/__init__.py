"""CPU oracle of the `latticenet` extension module the reference imports
(seq_lattice/lattice_modules.py:7-8, train_ln.py:16): HashTable, Lattice, ModelParams.

TEST INFRASTRUCTURE ONLY (part of oracle/).  PARITY UNPINNED -- see oracle/lattice_oracle.c.
Integer structure (keys, vertex ids, neighbour tables) and barycentric weights come from the scalar
C oracle; feature arithmetic is plain torch fp32 on the CPU so autograd supplies every backward.
"""
import os
import sys

import hjson
import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.abspath(os.path.join(_HERE, "..", "..", ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from oracle import lattice_oracle as O  # noqa: E402

POS_DIM = 3


class ModelParams:
    """Getter struct over the cfg `model` section (train_ln.py:80; getters used at
    seq_lattice/models.py:29-37,63-64,488,503)."""

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


class HashTable:
    """Thin view of the oracle's open-addressing table (imported, never used directly, by
    seq_lattice/lattice_modules.py:7)."""

    def __init__(self, capacity):
        self.table = O.OracleTable(capacity)

    def capacity(self): return self.table.capacity
    def nr_filled(self): return self.table.size()
    def keys(self): return self.table.keys()
    def clear(self): self.table.clear()


class Lattice:
    """Stateful lattice handle (`ls` in the reference).  One per resolution level; the coarser level
    hangs off `_coarse` and persists for the lifetime of the root (convention U3 / quirk Q2)."""

    def __init__(self, capacity, sigma, level=0, inv_std_dev=O.INV_STD_DEV_ADAMS):
        self.capacity = int(capacity)
        self.sigma = float(sigma)
        self.level = level
        self.inv_std_dev = inv_std_dev
        self.hash_table = HashTable(capacity)
        self._values = None
        self._positions = None
        self._frame = 0
        self._coarse = None
        self._coarse_frame = -1
        self._nbr_cache = {}

    # ---- reference-visible API -------------------------------------------------------------
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
        self._values = v

    def values(self): return self._values
    def val_dim(self): return int(self._values.shape[1])
    def pos_dim(self): return POS_DIM
    def positions(self): return self._positions
    def nr_lattice_vertices(self): return self.hash_table.nr_filled()

    def get_filter_extent(self, neighbourhood_size):
        if neighbourhood_size != 1:
            raise RuntimeError("only the 1-hop neighbourhood is supported (lattice_modules.py:299)")
        return 2 * (POS_DIM + 1) + 1

    # ---- structure ops -------------------------------------------------------------------------
    def scale(self):
        return O.scale_factors(self.sigma, self.inv_std_dev)

    def _bump(self):
        self._frame += 1
        self._nbr_cache = {}

    def distribute(self, positions, values, reset_hashmap=True):
        if reset_hashmap:
            self.hash_table.clear()
            self._coarse = None
        self._bump()
        self._positions = positions
        rows, idx, w = self.hash_table.table.distribute(positions.detach().numpy(), values.detach().numpy(), self.scale())
        return torch.from_numpy(rows), torch.from_numpy(idx), torch.from_numpy(w)

    def just_create_verts(self, positions, reset_hashmap=True):
        if reset_hashmap:
            self.hash_table.clear()
            self._coarse = None
        self._bump()
        self._positions = positions
        self.hash_table.table.insert_points(positions.detach().numpy(), self.scale())

    def create_coarse_verts(self):
        """Coarser lattice (sigma x2) holding every coarse vertex touched by any frame so far."""
        if self._coarse is None:
            self._coarse = Lattice(self.capacity, self.sigma * 2.0, self.level + 1, self.inv_std_dev)
        c = self._coarse
        if self._coarse_frame != self._frame:
            c.just_create_verts(self._positions, reset_hashmap=False)
            self._coarse_frame = self._frame
        return c

    def neighbours(self, other=None, mode=0, dilation=1):
        """[V,9] LongTensor of ids in `other` (default self); -1 absent; slot 8 = centre."""
        other = other or self
        key = (id(other), other._frame, mode, dilation)
        if key not in self._nbr_cache:
            n = self.hash_table.table.neighbours(other.hash_table.table, mode, dilation)
            self._nbr_cache[key] = torch.from_numpy(n.astype(np.int64))
        return self._nbr_cache[key]
