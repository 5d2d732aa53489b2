"""`latticenet_py.lattice.lattice_funcs` (seq_lattice/lattice_modules.py:14, models.py:6)."""
from temporal_latticenet_b200.funcs import (ConvIm2RowLattice, CoarsenLattice, DistributeLattice, FinefyLattice,  # noqa: F401
                                             GatherLattice, Im2RowIndicesLattice, Im2RowLattice, SliceClassifyLattice,
                                             SliceLattice, SplatLattice)
