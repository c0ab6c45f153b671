"""sonar-b200: B200-native fingerprint + alignment hot path of RyanBlaney/sonido-sonar.

The product is libsonar.so (hand-written sm_100a CUDA behind the C ABI in
include/sonar.h).  This package holds the kernels (csrc/), the ctypes binding
of that ABI (capi.py), the host-side mirror of the reference's Go API
(host/, C++ header-only), the multi-GPU partitioning helpers (sharding.py) and seeded synthetic inputs
(synth.py).  The directory name carries a hyphen (it is the reference's name);
import it with importlib.import_module("sonido-sonar_b200").
"""

from . import capi, sharding, synth  # noqa: E402,F401

__all__ = ["capi", "sharding", "synth"]
