"""`latticenet_py.lattice.lovasz_loss` (train_ln.py:17)."""
from temporal_latticenet_b200.lovasz import LovaszSoftmax  # noqa: F401
