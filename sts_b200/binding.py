"""ctypes view of include/stsgpu.h.  No compute, no fallback: every call goes to libstsgpu.so and a
missing library or device raises EngineError."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = os.environ.get("STSGPU_LIB") or os.path.join(_HERE, "libstsgpu.so")	# override: A/B builds of the engine

ALL_TESTS = 0xFFFE
NON_P_VALUE = -99999999.0
ABI_VERSION = 2
FLAG_STATS = 1
UNIQUE_ID_BYTES = 128
TEST_NAMES = [None, "Frequency", "BlockFrequency", "CumulativeSums", "Runs", "LongestRun", "Rank", "DFT",
              "NonOverlappingTemplate", "OverlappingTemplate", "Universal", "ApproximateEntropy",
              "RandomExcursions", "RandomExcursionsVariant", "Serial", "LinearComplexity"]

# every symbol include/stsgpu.h declares (checked by tests/test_abi.py)
EXPORTS = ["stsgpu_default_params", "stsgpu_create", "stsgpu_destroy", "stsgpu_get_layout", "stsgpu_query_layout", "stsgpu_dft_supported", "stsgpu_submit",
           "stsgpu_submit_resident", "stsgpu_submit_ascii", "stsgpu_ascii_status", "stsgpu_upload", "stsgpu_wait", "stsgpu_pending", "stsgpu_results", "stsgpu_results_stats", "stsgpu_generate_lcg", "stsgpu_generate",
           "stsgpu_download_input", "stsgpu_host_alloc", "stsgpu_host_free", "stsgpu_comm_unique_id",
           "stsgpu_comm_init", "stsgpu_gather_begin", "stsgpu_gather_pvals", "stsgpu_comm_barrier", "stsgpu_assess_begin",
           "stsgpu_assess_add", "stsgpu_assess_reduce", "stsgpu_assess_finish", "stsgpu_set_profiling",
           "stsgpu_kernel_count", "stsgpu_kernel_info", "stsgpu_last_batch_ms", "stsgpu_launch_count",
           "stsgpu_debug_pattern_histogram", "stsgpu_strerror", "stsgpu_last_error", "stsgpu_abi_version"]


class EngineError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n", C.c_int64), ("max_streams", C.c_int64),
                ("test_mask", C.c_uint32), ("pipeline_depth", C.c_int32), ("block_frequency_M", C.c_int64),
                ("non_overlapping_m", C.c_int32), ("overlapping_m", C.c_int32), ("apen_m", C.c_int32),
                ("serial_m", C.c_int32), ("linear_complexity_M", C.c_int64), ("flags", C.c_uint32),
                ("reserved", C.c_uint32)]


class Metric(C.Structure):
    _fields_ = [("sample_count", C.c_int64), ("too_low", C.c_int64), ("pass_count", C.c_int64),
                ("proportion_min", C.c_double), ("proportion_max", C.c_double), ("chi2", C.c_double),
                ("uniformity", C.c_double), ("uniformity_passed", C.c_int32), ("proportion_passed", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [("enabled_mask", C.c_uint32), ("npv", C.c_int32), ("nst", C.c_int32), ("nw", C.c_int32),
                ("ntemplates", C.c_int32), ("universal_L", C.c_int32),
                ("pv_off", C.c_int32 * 16), ("pv_cnt", C.c_int32 * 16),
                ("st_off", C.c_int32 * 16), ("st_cnt", C.c_int32 * 16),
                ("nfs", C.c_int32), ("fs_off", C.c_int32 * 16), ("fs_cnt", C.c_int32 * 16)]


def library_path():
    return _LIBNAME


def build_library(force=False):
    """compile sts_b200/libstsgpu.so for sm_100a with nvcc (cross-compiles without a GPU)"""
    if force:
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean"])
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIBNAME


_lib = None
_ALLOW_EMULATION = False	# only tests/emu_util.py sets this, on its private copy of this module


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBNAME):
        raise EngineError("%s is missing: run `make -C sts_b200` (there is no CPU fallback)" % _LIBNAME)
    lib = C.CDLL(_LIBNAME)
    if hasattr(lib, "stsgpu_is_emulation") and not _ALLOW_EMULATION:
        raise EngineError("%s is the CPU emulation of tests/emu - test infrastructure, never a fallback: the product "
                          "only runs on the CUDA build of sts_b200/libstsgpu.so" % _LIBNAME)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.stsgpu_default_params.argtypes = [C.POINTER(Params)]
    lib.stsgpu_default_params.restype = None
    lib.stsgpu_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.stsgpu_destroy.argtypes = [vp]
    lib.stsgpu_destroy.restype = None
    lib.stsgpu_get_layout.argtypes = [vp, C.POINTER(Layout)]
    lib.stsgpu_query_layout.argtypes = [C.POINTER(Params), C.POINTER(Layout)]
    lib.stsgpu_dft_supported.argtypes = [i64]
    lib.stsgpu_submit.argtypes = [vp, vp, i64, i64]
    lib.stsgpu_submit_resident.argtypes = [vp, i64, i64]
    lib.stsgpu_submit_ascii.argtypes = [vp, vp, i64, i64, i64]
    lib.stsgpu_ascii_status.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.stsgpu_upload.argtypes = [vp, vp, i64]
    lib.stsgpu_wait.argtypes = [vp]
    lib.stsgpu_pending.argtypes = [vp]
    lib.stsgpu_results.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.stsgpu_results_stats.argtypes = [vp, C.POINTER(vp)]
    lib.stsgpu_generate_lcg.argtypes = [vp, i64, i64]
    lib.stsgpu_generate.argtypes = [vp, C.c_int, i64, i64]
    lib.stsgpu_download_input.argtypes = [vp, vp, i64]
    lib.stsgpu_host_alloc.argtypes = [C.c_size_t]
    lib.stsgpu_host_alloc.restype = vp
    lib.stsgpu_host_free.argtypes = [vp]
    lib.stsgpu_host_free.restype = None
    lib.stsgpu_comm_unique_id.argtypes = [vp]
    lib.stsgpu_comm_init.argtypes = [vp, vp, i32, i32]
    lib.stsgpu_gather_begin.argtypes = [vp, i32]
    lib.stsgpu_gather_pvals.argtypes = [vp, i32, C.POINTER(vp)]
    lib.stsgpu_comm_barrier.argtypes = [vp]
    lib.stsgpu_assess_begin.argtypes = [vp, i64, C.c_double, C.c_double]
    lib.stsgpu_assess_add.argtypes = [vp, i32, vp, i64, i64]
    lib.stsgpu_assess_reduce.argtypes = [vp]
    lib.stsgpu_assess_finish.argtypes = [vp, C.POINTER(Metric), vp]
    lib.stsgpu_set_profiling.argtypes = [vp, i32]
    lib.stsgpu_kernel_count.argtypes = [vp]
    lib.stsgpu_kernel_info.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    lib.stsgpu_last_batch_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.stsgpu_launch_count.argtypes = [vp]
    lib.stsgpu_launch_count.restype = i64
    lib.stsgpu_debug_pattern_histogram.argtypes = [vp, i64, i32, vp]
    lib.stsgpu_strerror.argtypes = [i32]
    lib.stsgpu_strerror.restype = C.c_char_p
    lib.stsgpu_last_error.argtypes = [vp]
    lib.stsgpu_last_error.restype = C.c_char_p
    lib.stsgpu_abi_version.restype = i32
    _lib = lib
    return lib


def dft_supported(n):
    """can the engine's FFT take streams of n bits (any even n up to 2^25; beyond that n/2 = two 5-smooth factors <= 4096)"""
    return bool(load_library().stsgpu_dft_supported(n))


def query_layout(params=None, **kw):
    """record layout for a parameter set without an engine or a device (stsgpu_query_layout)"""
    p = params if params is not None else default_params(**kw)
    lay = Layout()
    rc = load_library().stsgpu_query_layout(C.byref(p), C.byref(lay))
    if rc != 0:
        raise EngineError("stsgpu_query_layout: %s" % load_library().stsgpu_strerror(rc).decode())
    return lay


def default_params(**kw):
    p = Params()
    load_library().stsgpu_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown parameter %s" % k)
        setattr(p, k, v)
    return p


class PinnedBuffer:
    """page-locked host bytes from stsgpu_host_alloc, exposed as a numpy uint8 array"""

    def __init__(self, nbytes):
        self._lib = load_library()
        self.ptr = self._lib.stsgpu_host_alloc(nbytes)
        if not self.ptr:
            raise EngineError("stsgpu_host_alloc(%d) failed" % nbytes)
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def free(self):
        if self.ptr:
            self._lib.stsgpu_host_free(self.ptr)
            self.ptr = None
            self.array = None


class Engine:
    """one engine context (one GPU).  Mirrors the C call sequence: submit -> wait -> results."""

    def __init__(self, params=None, **kw):
        self.lib = load_library()
        self.params = params if params is not None else default_params(**kw)
        h = C.c_void_p()
        rc = self.lib.stsgpu_create(C.byref(self.params), C.byref(h))
        if rc != 0:
            raise EngineError("stsgpu_create: %s (%s)" % (self.lib.stsgpu_strerror(rc).decode(),
                                                         self.lib.stsgpu_last_error(None).decode()))
        self.h = h
        self.layout = Layout()
        self._ck(self.lib.stsgpu_get_layout(self.h, C.byref(self.layout)), "get_layout")

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError("%s: %s (%s)" % (what, self.lib.stsgpu_strerror(rc).decode(),
                                               self.lib.stsgpu_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.stsgpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the iterate() replacement ----
    def submit(self, raw, nstreams, first_iteration=0):
        if isinstance(raw, np.ndarray):
            assert raw.dtype == np.uint8 and raw.flags["C_CONTIGUOUS"]
            assert raw.size >= nstreams * self.params.n // 8
            ptr = raw.ctypes.data
            self._keep = (getattr(self, "_keep", ()) + (raw,))[-4:]	# borrowed until the batch's wait
        else:
            ptr = raw
        self._ck(self.lib.stsgpu_submit(self.h, ptr, nstreams, first_iteration), "submit")

    def submit_ascii(self, text, nstreams, first_iteration=0):
        """text: bytes-like / uint8 array of '0' / '1' characters and white space, starting at the batch's first stream"""
        text = np.ascontiguousarray(np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text)
        self._keep = (getattr(self, "_keep", ()) + (text,))[-4:]
        self._ck(self.lib.stsgpu_submit_ascii(self.h, text.ctypes.data, text.size, nstreams, first_iteration), "submit_ascii")

    def ascii_status(self):
        done, bad = C.c_int64(), C.c_int64()
        self._ck(self.lib.stsgpu_ascii_status(self.h, C.byref(done), C.byref(bad)), "ascii_status")
        return done.value, bad.value

    def submit_resident(self, nstreams, first_iteration=0):
        self._ck(self.lib.stsgpu_submit_resident(self.h, nstreams, first_iteration), "submit_resident")

    def upload(self, raw, nstreams):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        self._keep = (getattr(self, "_keep", ()) + (raw,))[-4:]
        self._ck(self.lib.stsgpu_upload(self.h, raw.ctypes.data, nstreams), "upload")

    def wait(self):
        self._ck(self.lib.stsgpu_wait(self.h), "wait")

    def pending(self):
        return self.lib.stsgpu_pending(self.h)

    def results(self, copy=True):
        ns, first = C.c_int64(), C.c_int64()
        pv, st, w = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.lib.stsgpu_results(self.h, C.byref(ns), C.byref(first), C.byref(pv), C.byref(st),
                                         C.byref(w)), "results")
        lay, k = self.layout, ns.value
        pva = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(k, lay.npv))
        sta = np.ctypeslib.as_array(C.cast(st, C.POINTER(C.c_int64)), shape=(k, lay.nst))
        if lay.nw:
            wa = np.ctypeslib.as_array(C.cast(w, C.POINTER(C.c_uint32)), shape=(k, lay.nw))
        else:
            wa = np.zeros((k, 0), dtype=np.uint32)
        if copy:
            pva, sta, wa = pva.copy(), sta.copy(), wa.copy()
        return pva, sta, wa

    def results_stats(self):
        """fstats[nstreams, nfs] of the last completed batch (engines created with flags=FLAG_STATS), else None"""
        ns, fs = C.c_int64(), C.c_void_p()
        self._ck(self.lib.stsgpu_results(self.h, C.byref(ns), None, None, None, None), "results")
        self._ck(self.lib.stsgpu_results_stats(self.h, C.byref(fs)), "results_stats")
        if not fs.value:
            return None
        return np.ctypeslib.as_array(C.cast(fs, C.POINTER(C.c_double)), shape=(ns.value, self.layout.nfs)).copy()

    def run(self, raw, nstreams, first_iteration=0):
        self.submit(raw, nstreams, first_iteration)
        self.wait()
        return self.results()

    # ---- synthetic input ----
    def generate_lcg(self, first_stream, nstreams):
        self._ck(self.lib.stsgpu_generate_lcg(self.h, first_stream, nstreams), "generate_lcg")

    def generate(self, generator, first_stream, nstreams):
        """generators 1, 2, 4, 5 of tools/generators.c into the device input buffer (stsgpu_generate)"""
        self._ck(self.lib.stsgpu_generate(self.h, generator, first_stream, nstreams), "generate")

    def download_input(self, nstreams):
        out = np.zeros(nstreams * self.params.n // 8, dtype=np.uint8)
        self._ck(self.lib.stsgpu_download_input(self.h, out.ctypes.data, nstreams), "download_input")
        return out

    # ---- multi GPU ----
    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
        rc = load_library().stsgpu_comm_unique_id(buf)
        if rc != 0:
            raise EngineError("stsgpu_comm_unique_id failed: %d" % rc)
        return bytes(buf)

    def comm_init(self, uid, rank, nranks):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(uid)
        self._ck(self.lib.stsgpu_comm_init(self.h, buf, rank, nranks), "comm_init")

    def gather_begin(self, root=0):
        """gather mode: every batch submitted from now on ends with the exchange of its p-value records to `root`"""
        self._ck(self.lib.stsgpu_gather_begin(self.h, root), "gather_begin")

    def gather_pvals(self, root=0, nranks=1, copy=True):
        """p-values of the last completed batch of every rank, rank-major, on `root` (None elsewhere); copy=False
        returns a view of the engine's pinned buffer, valid until that batch slot completes again"""
        out = C.c_void_p()
        self._ck(self.lib.stsgpu_gather_pvals(self.h, root, C.byref(out)), "gather_pvals")
        if not out.value:
            return None
        ns = C.c_int64()
        self.lib.stsgpu_results(self.h, C.byref(ns), None, None, None, None)
        a = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(nranks * ns.value, self.layout.npv))
        return a.copy() if copy else a

    def comm_barrier(self):
        self._ck(self.lib.stsgpu_comm_barrier(self.h), "comm_barrier")

    # ---- assessment tallies on the device ----
    def assess_begin(self, bins, alpha=0.01, uniformity_level=0.0001):
        self._assess_bins = bins
        self._ck(self.lib.stsgpu_assess_begin(self.h, bins, alpha, uniformity_level), "assess_begin")

    def assess_add(self, test, pvals, first_index=0):
        """p-values of one test in .pvalues order (iteration-major) into the tallies"""
        v = np.ascontiguousarray(pvals, dtype=np.float64).reshape(-1)
        self._ck(self.lib.stsgpu_assess_add(self.h, test, v.ctypes.data, v.size, first_index), "assess_add")

    def assess_reduce(self):
        self._ck(self.lib.stsgpu_assess_reduce(self.h), "assess_reduce")

    def assess_finish(self):
        """-> dict of per-column arrays (one entry per p-value column, record order)"""
        npv, bins = self.layout.npv, self._assess_bins
        m = (Metric * npv)()
        freq = np.zeros((npv, bins), dtype=np.int64)
        self._ck(self.lib.stsgpu_assess_finish(self.h, m, freq.ctypes.data), "assess_finish")
        g = lambda f, dt: np.array([getattr(x, f) for x in m], dtype=dt)
        return dict(sample=g("sample_count", np.int64), toolow=g("too_low", np.int64), passed=g("pass_count", np.int64),
                    freq=freq, tmin=g("proportion_min", np.float64), tmax=g("proportion_max", np.float64),
                    chi2=g("chi2", np.float64), uniformity=g("uniformity", np.float64),
                    flags=g("uniformity_passed", np.int32) | (g("proportion_passed", np.int32) << 1))

    # ---- measurement ----
    def set_profiling(self, on):
        self._ck(self.lib.stsgpu_set_profiling(self.h, 1 if on else 0), "set_profiling")

    def kernel_times(self):
        out = []
        for i in range(self.lib.stsgpu_kernel_count(self.h)):
            name, ms = C.c_char_p(), C.c_float()
            self._ck(self.lib.stsgpu_kernel_info(self.h, i, C.byref(name), C.byref(ms)), "kernel_info")
            out.append((name.value.decode(), ms.value))
        return out

    def last_batch_ms(self):
        ms = C.c_float()
        self._ck(self.lib.stsgpu_last_batch_ms(self.h, C.byref(ms)), "last_batch_ms")
        return ms.value

    def launch_count(self):
        return self.lib.stsgpu_launch_count(self.h)

    def debug_pattern_histogram(self, s, bits):
        h = np.zeros(1 << bits, dtype=np.uint32)
        self._ck(self.lib.stsgpu_debug_pattern_histogram(self.h, s, bits, h.ctypes.data), "debug_pattern_histogram")
        return h
