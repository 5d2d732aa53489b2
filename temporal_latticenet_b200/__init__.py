"""temporal_latticenet_b200 -- Blackwell-native (sm_100a) permutohedral-lattice hot path behind the
operator API that AIS-Bonn/temporal_latticenet's seq_lattice/models.py drives.

The CUDA extension (csrc/libltn_b200.so) is loaded on first use and there is NO CPU fallback:
every op raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
