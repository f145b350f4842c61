"""CPU: the bodies of tests/test_integration.py with integration/glue.c + the reference's own host translation units
linked against the emulated engine of tests/emu instead of libstsgpu.so (tests/emu/_build/sts_gpu_emu, built by
integration/Makefile with ENGINE_DIR / ENGINE_NAME pointing there).  Checks the glue's logic without a GPU; same caveat
as tests/test_emu_parity.py: it is not the product's parity claim."""
import os
import subprocess

import pytest

import test_integration as I
from emu_util import emu_library
from test_gpu_host_driver import lcg8  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.environ.get("STS_EMU_FULL", "") not in ("", "0")
slow = pytest.mark.skipif(not FULL, reason="slow on the emulator: set STS_EMU_FULL=1")


def emu_sts_gpu():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("the reference tree is needed to build the drop-in")
    lib = emu_library()
    build = os.path.dirname(lib)
    exe = os.path.join(build, "sts_gpu_emu")
    deps = [lib, os.path.join(ROOT, "integration", "glue.c"), os.path.join(ROOT, "integration", "Makefile"),
            os.path.join(ROOT, "include", "stsgpu.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "integration"), "ENGINE_DIR=" + build,
                               "ENGINE_NAME=stsgpu_emu", "OUT=" + os.path.join(build, "integration_obj"), "EXE=" + exe])
    return exe


@pytest.fixture(autouse=True)
def emulated(monkeypatch):
    monkeypatch.setattr(I, "STS_GPU", emu_sts_gpu())


def _again(fn):
    import functools
    import types
    g = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    g = functools.update_wrapper(g, fn)
    del g.__wrapped__
    return g


test_no_cpu_test_code_is_linked = _again(I.test_no_cpu_test_code_is_linked)
test_result_txt_and_pvalues_against_reference = slow(_again(I.test_result_txt_and_pvalues_against_reference))
test_iterate_only_jobs_then_assess_only = _again(I.test_iterate_only_jobs_then_assess_only)
test_dash_s_stats_results_and_data_files_identical = _again(I.test_dash_s_stats_results_and_data_files_identical)
test_dash_s_on_degenerate_streams = _again(I.test_dash_s_on_degenerate_streams)
test_non_default_parameters_and_legacy_output = _again(I.test_non_default_parameters_and_legacy_output)
test_ascii_and_stdin_input = slow(_again(I.test_ascii_and_stdin_input))
test_short_raw_file_is_fatal_like_the_reference = _again(I.test_short_raw_file_is_fatal_like_the_reference)
