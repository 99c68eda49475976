"""alis_b200 -- B200-native FDTD time-stepping path of ALiS behind the reference's C API.

The product is the CUDA engine (alis_b200/csrc, sm_100a) reached through the C ABI in
include/alis_b200.h, plus the ALiS-compatible host C layer in alis_b200/host.  This Python
package only holds the ctypes binding used by bench.py / tests and the Python mirror of the
reference setup for synthetic benchmark worlds.  Importing it does not load any library;
`capi.load_engine()` fails loudly when the CUDA extension has not been built.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
