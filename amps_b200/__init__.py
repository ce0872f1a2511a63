"""amps_b200 -- B200-native charged-particle push + ECSIM J/mass-matrix deposition for AMPS.

The product is the C-ABI CUDA library ``libamps_gpu.so`` (include/amps_gpu.h); this package holds
its sources (csrc/), the host-side mesh flattening (mesh.py) and a thin ctypes mirror of the
reference entry points (pic.py).  There is no CPU fallback.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
