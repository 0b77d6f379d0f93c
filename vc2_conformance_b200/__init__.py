"""
vc2_conformance_b200 -- B200-native (sm_100a) implementation of the data-parallel picture path
of SMPTE ST 2042-1 (VC-2), drop-in for the corresponding functions of bbc/vc2_conformance.

Layers:
  csrc/           hand-written CUDA kernels + the C ABI (include/vc2_b200.h) -> _lib/libvc2b200.so
  capi.py         ctypes binding (no CPU fallback: raises if the library is missing)
  params.py       picture-path parameters (the subset of the reference State the kernels read)
  device.py       device-resident batch decoder/encoder (throughput API)
  pseudocode.py   reference-named functions (idwt, dwt, inverse_quant, picture_decode, ...)
  install.py      rebinding of the reference's call sites
"""

__version__ = "0.1.0"

from . import capi  # noqa: F401
from .params import make_params, params_from_state, Geometry  # noqa: F401
