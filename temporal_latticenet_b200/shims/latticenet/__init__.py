"""`latticenet` as the reference imports it (seq_lattice/lattice_modules.py:7-8, train_ln.py:16),
backed by temporal_latticenet_b200 (sm_100a kernels; no CPU fallback)."""
from temporal_latticenet_b200.lattice import HashTable, Lattice, ModelParams  # noqa: F401
