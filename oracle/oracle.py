"""ctypes binding of oracle/libsts_oracle.so (see oracle/sts_oracle.h).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsts_oracle.so")

TEST_NAMES = [None, "Frequency", "BlockFrequency", "CumulativeSums", "Runs", "LongestRun", "Rank", "DFT",
              "NonOverlappingTemplate", "OverlappingTemplate", "Universal", "ApproximateEntropy",
              "RandomExcursions", "RandomExcursionsVariant", "Serial", "LinearComplexity"]
ALL_TESTS = 0xFFFE
NON_P_VALUE = -99999999.0


class Params(C.Structure):
    _fields_ = [("n", C.c_int64), ("test_mask", C.c_uint32), ("flags", C.c_int32),
                ("block_frequency_M", C.c_int64), ("non_overlapping_m", C.c_int32),
                ("overlapping_m", C.c_int32), ("apen_m", C.c_int32), ("serial_m", C.c_int32),
                ("linear_complexity_M", C.c_int64)]


class Layout(C.Structure):
    _fields_ = [("enabled_mask", C.c_uint32), ("npv", C.c_int32), ("nst", C.c_int32), ("nw", C.c_int32),
                ("ntemplates", C.c_int32), ("universal_L", C.c_int32),
                ("pv_off", C.c_int32 * 16), ("pv_cnt", C.c_int32 * 16),
                ("st_off", C.c_int32 * 16), ("st_cnt", C.c_int32 * 16)]


def build(force=False):
    """compile the C restatement (gcc only); also oracle/_ref when /root/reference is present"""
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_layout_for.argtypes = [C.POINTER(Params), C.POINTER(Layout)]
        _lib.oracle_run.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int]
        _lib.oracle_pattern_histogram.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        _lib.oracle_lcg_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        _lib.oracle_lcg_bytes.restype = None
        _lib.oracle_assess.argtypes = [C.POINTER(Layout), C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_igamc.argtypes = [C.c_double, C.c_double]
        _lib.oracle_igamc.restype = C.c_double
        _lib.oracle_igam.argtypes = [C.c_double, C.c_double]
        _lib.oracle_igam.restype = C.c_double
    return _lib


def make_params(n=1 << 20, test_mask=ALL_TESTS, block_frequency_M=16384, non_overlapping_m=9,
                overlapping_m=9, apen_m=10, serial_m=16, linear_complexity_M=500, flags=0):
    return Params(n, test_mask, flags, block_frequency_M, non_overlapping_m, overlapping_m, apen_m, serial_m,
                  linear_complexity_M)


def layout(params):
    lay = Layout()
    lib().oracle_layout_for(C.byref(params), C.byref(lay))
    return lay


def run(params, raw, nstreams, nthreads=1):
    """raw: uint8 array of nstreams*n/8 file bytes -> (pvals[ns,npv], istats[ns,nst], wcount[ns,nw])"""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    assert raw.size >= nstreams * params.n // 8
    lay = layout(params)
    pv = np.zeros((nstreams, lay.npv), dtype=np.float64)
    st = np.zeros((nstreams, lay.nst), dtype=np.int64)
    w = np.zeros((nstreams, max(lay.nw, 1)), dtype=np.uint32)
    rc = lib().oracle_run(C.byref(params), raw.ctypes.data, nstreams, pv.ctypes.data, st.ctypes.data,
                          w.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("oracle_run failed: %d" % rc)
    return pv, st, w[:, :lay.nw], lay


def pattern_histogram(raw, n, bits):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    h = np.zeros(1 << bits, dtype=np.uint32)
    lib().oracle_pattern_histogram(raw.ctypes.data, n, bits, h.ctypes.data)
    return h


def lcg_bytes(first_stream, nstreams, n=1 << 20):
    out = np.zeros(nstreams * n // 8, dtype=np.uint8)
    lib().oracle_lcg_bytes(first_stream, nstreams, n, out.ctypes.data)
    return out


def assess(lay, pv, bins, alpha=0.01, uniformity_level=0.0001):
    """uniformity / proportion assessment of pv[iterations, npv] -> dict of per-column arrays"""
    pv = np.ascontiguousarray(pv, dtype=np.float64)
    it, npv = pv.shape
    tally = np.zeros((npv, 3), dtype=np.int64)
    freq = np.zeros((npv, bins), dtype=np.int64)
    fp = np.zeros((npv, 4), dtype=np.float64)
    flags = np.zeros(npv, dtype=np.int32)
    rc = lib().oracle_assess(C.byref(lay), pv.ctypes.data, it, bins, alpha, uniformity_level, tally.ctypes.data,
                             freq.ctypes.data, fp.ctypes.data, flags.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_assess failed: %d" % rc)
    return dict(sample=tally[:, 0], toolow=tally[:, 1], passed=tally[:, 2], freq=freq, tmin=fp[:, 0], tmax=fp[:, 1],
                chi2=fp[:, 2], uniformity=fp[:, 3], flags=flags)


def igamc(a, x):
    return lib().oracle_igamc(a, x)
