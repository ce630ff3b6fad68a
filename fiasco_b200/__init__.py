"""fiasco_b200 -- B200-native FIASCO encoder hot path.

The product is native: `csrc/` (hand-written sm_100a CUDA + the C ABI of
include/fiasco_b200.h) and `host/` (the libfiasco C API: fiasco_coder(),
fiasco_c_options_*).  This Python package is only the ctypes binding used by tests and
bench.py; it never computes anything itself and raises if the CUDA library is missing.
"""
from .ffi import (  # noqa: F401
    FB200Error,
    Motion,
    Params,
    TileEncoder,
    device_count,
    lib_path,
    load,
    pixels_from_grey,
    probe,
    wfa_lines,
)
