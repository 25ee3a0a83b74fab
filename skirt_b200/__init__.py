"""skirt_b200 -- B200-native photon-packet engine for SKIRT's propagation hot path.

This package is only the Python face of the C ABI in include/skirtgpu.h (libskirtgpu.so, CUDA
sm_100a).  There is no CPU fallback: importing works anywhere, but creating an Engine raises
unless the shared library is built and a CUDA device is present.
"""
from .binding import Engine, EngineError, LIB_PATH, load_library, lib_available  # noqa: F401
