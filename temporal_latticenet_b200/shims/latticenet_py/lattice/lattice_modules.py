"""`latticenet_py.lattice.lattice_modules` (seq_lattice/lattice_modules.py:15, models.py:7)."""
from temporal_latticenet_b200.modules import (BottleneckBlock, Conv1x1, ConvLatticeModule, CoarsenLatticeModule,  # noqa: F401
                                               DistributeLatticeModule, FinefyLatticeModule, Gn, GnRelu1x1, GnReluCoarsen,
                                               GnReluConv, GnReluFinefy, GroupNormLatticeModule, ResnetBlock,
                                               SliceFastCUDALatticeModule, SliceLatticeModule, SplatLatticeModule)
