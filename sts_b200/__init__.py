"""sts_b200 - B200-native engine for the per-bitstream test kernels of arcetri/sts.

The product is the C-ABI shared library `sts_b200/libstsgpu.so` (include/stsgpu.h) plus the C host
driver under sts_b200/host.  This Python package is only a thin ctypes view of that ABI for the test
suite and bench.py; it contains no compute and never imports the oracle.
"""
from .binding import (dft_supported, query_layout, ALL_TESTS, FLAG_STATS, NON_P_VALUE, TEST_NAMES, Engine, EngineError, Layout, Params, build_library,
                      default_params, library_path, load_library)

__all__ = ["dft_supported", "query_layout", "ALL_TESTS", "FLAG_STATS", "NON_P_VALUE", "TEST_NAMES", "Engine", "EngineError", "Layout", "Params", "build_library",
           "default_params", "library_path", "load_library"]
