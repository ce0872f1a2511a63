"""amps_b200 -- B200-native charged-particle push + ECSIM J/mass-matrix deposition for AMPS.

The product is the C-ABI CUDA library ``libamps_gpu.so`` (include/amps_gpu.h); this package holds
its sources (csrc/), the C++ host layer on the reference's own data structures (host/amps_gpu_host.hpp, host/amps_gpu_host_mesh.hpp),
the Python builder of flattened mesh descriptions for synthetic boxes (mesh.py), the synthetic workloads of the benchmark
(workload.py) and a thin ctypes mirror of the reference entry points (api.py over _capi.py).  There is no CPU fallback.
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
