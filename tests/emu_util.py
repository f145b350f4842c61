"""TEST INFRASTRUCTURE: the engine's UNMODIFIED CUDA sources compiled for the CPU emulator of tests/emu (see
tests/emu/cuda_emu.h) and loaded through a private copy of the ctypes binding.  Used by the `-m "not gpu"` suite to
step the real kernels' logic against the oracle and the golden vectors where no GPU exists; the product package
sts_b200 never sees this library."""
import importlib.util
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_mod = None


def emu_library():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    return build_emu.build()


def emu_binding():
    """a second instance of sts_b200/binding.py bound to tests/emu/_build/libstsgpu_emu.so"""
    global _mod
    if _mod is None:
        lib = emu_library()
        spec = importlib.util.spec_from_file_location("sts_b200_emu_binding", os.path.join(ROOT, "sts_b200", "binding.py"))
        mod = importlib.util.module_from_spec(spec)
        old = os.environ.get("STSGPU_LIB")
        os.environ["STSGPU_LIB"] = lib
        try:
            spec.loader.exec_module(mod)
        finally:
            if old is None:
                del os.environ["STSGPU_LIB"]
            else:
                os.environ["STSGPU_LIB"] = old
        assert mod.library_path() == lib
        mod._ALLOW_EMULATION = True		# the product's own instance of the binding refuses this library
        mod.load_library()
        _mod = mod
    return _mod


def emu_host_driver():
    """sts_b200/host (the C host driver) linked against the emulated engine: tests/emu/_build/stsb200_emu"""
    lib = emu_library()
    build = os.path.dirname(lib)
    exe = os.path.join(build, "stsb200_emu")
    host = os.path.join(ROOT, "sts_b200", "host")
    srcs = [os.path.join(host, "stsb200.c"), os.path.join(host, "sts_host.c"), os.path.join(host, "sts_generators.c")]
    deps = srcs + [os.path.join(host, "sts_host.h"), lib]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        asan = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if os.environ.get("STS_EMU_ASAN", "") not in ("", "0") else []
        subprocess.check_call([os.environ.get("CC", "gcc"), "-std=c99", "-O2", "-Wall", "-I", os.path.join(ROOT, "include")] + asan + srcs +
                              ["-L", build, "-lstsgpu_emu", "-Wl,-rpath," + build, "-lpthread", "-lm", "-o", exe])
    return exe
