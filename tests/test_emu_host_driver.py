"""CPU: the C host driver tests of tests/test_gpu_host_driver.py re-run with sts_b200/host (stsb200.c + sts_host.c,
unmodified) linked against the emulated engine of tests/emu instead of libstsgpu.so: the whole host logic - batching
loop with ragged batches, -j seek rule, -F a, stdin, record keeping, .pvalues writer and bounded-memory reader,
result.txt / finalAnalysisReport.txt / -s files - against the reference binary and the committed golden bodies,
without a GPU.  Same caveat as tests/test_emu_parity.py: this checks logic, it is not the product's parity claim."""
import os

import pytest

import test_gpu_host_driver as H
from emu_util import emu_host_driver
from test_gpu_host_driver import lcg8  # noqa: F401  (module-scoped fixture: 8 LCG streams in a file)

FULL = os.environ.get("STS_EMU_FULL", "") not in ("", "0")
slow = pytest.mark.skipif(not FULL, reason="slow on the emulator: set STS_EMU_FULL=1")


@pytest.fixture(autouse=True)
def emulated_driver(monkeypatch):
    monkeypatch.setattr(H, "STSB200", emu_host_driver())


def _again(fn):
    import functools
    import types
    g = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    g = functools.update_wrapper(g, fn)
    del g.__wrapped__
    return g


test_pvalues_file_matches_reference_golden = _again(H.test_pvalues_file_matches_reference_golden)
test_result_txt_identical_to_reference = _again(H.test_result_txt_identical_to_reference)
test_legacy_assessment_table_identical_to_reference = _again(H.test_legacy_assessment_table_identical_to_reference)
test_jobnum_and_ascii_input = _again(H.test_jobnum_and_ascii_input)
test_short_file_is_fatal = _again(H.test_short_file_is_fatal)
test_ascii_file_with_line_breaks_matches_reference = slow(_again(H.test_ascii_file_with_line_breaks_matches_reference))
test_stdin_input = _again(H.test_stdin_input)
test_result_txt_of_mode_b_identical_to_reference = _again(H.test_result_txt_of_mode_b_identical_to_reference)
test_assess_only_mode_streams_pvalues_files = _again(H.test_assess_only_mode_streams_pvalues_files)
test_dash_s_results_and_data_files_identical_to_reference = _again(H.test_dash_s_results_and_data_files_identical_to_reference)
test_generator_1_on_the_device = _again(H.test_generator_1_on_the_device)
test_ascii_file_against_committed_reference_reading = _again(H.test_ascii_file_against_committed_reference_reading)
test_reports_against_committed_reference_bodies = _again(H.test_reports_against_committed_reference_bodies)
test_generators_2_to_9_are_the_files_the_tool_writes = _again(H.test_generators_2_to_9_are_the_files_the_tool_writes)
