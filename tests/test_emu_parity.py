"""CPU: the parity tests of tests/test_gpu_parity.py re-run with the engine's UNMODIFIED CUDA sources compiled for
the functional emulator of tests/emu (fibers for the threads of a CTA, real warp collectives and barriers, a
synchronous runtime).  Same test bodies, same bars (integers bit-exact, p-values within 1e-9 of the reference's
golden values and of the oracle); only the `sts` fixture differs.

What this does and does not show: the kernels' LOGIC - index arithmetic, bit extraction, the warp-collective and
shared-memory algorithms, record layout, every host-side decision of engine.cu - is checked where no GPU exists,
on every round's CPU run.  It is not a parity claim for the product (that is `-m gpu` on a B200: nvcc's code
generation, libdevice, real concurrency and memory ordering are not emulated), nothing is timed on it, and the
product never loads this library (tests/test_abi.py::test_product_never_imports_the_oracle, and
sts_b200/binding.py refuses to run without its CUDA library).

The slowest cases (tens of seconds each on the emulator) only run with STS_EMU_FULL=1."""
import os

import pytest

import test_gpu_parity as G
from emu_util import emu_binding

FULL = os.environ.get("STS_EMU_FULL", "") not in ("", "0")
slow = pytest.mark.skipif(not FULL, reason="slow on the emulator: set STS_EMU_FULL=1")


@pytest.fixture(scope="module")
def sts():
    return emu_binding()


def _pick(fn, keep):
    """the parametrized test `fn` restricted to the parameter sets `keep` accepts (all of them with STS_EMU_FULL=1)"""
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    assert len(marks) == 1
    names, values = marks[0].args[0], marks[0].args[1]
    chosen = [v for v in values if FULL or (keep(*v) if isinstance(v, tuple) else keep(v))]

    return pytest.mark.parametrize(names, chosen)(_rebind(fn, names))


def _rebind(fn, names):
    import functools
    import types
    g = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    g = functools.update_wrapper(g, fn)
    g.pytestmark = [m for m in getattr(fn, "pytestmark", []) if m.name != "parametrize"]
    del g.__wrapped__
    return g


# the golden vectors of the unmodified reference binary and the oracle, all 15 tests
test_engine_matches_reference_golden_and_oracle = _pick(G.test_engine_matches_reference_golden_and_oracle,
                                                        lambda case: case != "patterns_6")
test_engine_config4_parameters = slow(_rebind(G.test_engine_config4_parameters, ""))
@slow
def test_full_batch_properties_and_sampled_parity(sts, oracle, monkeypatch):     # 1024 streams: ~2.5 min
    monkeypatch.setattr(G, "FULL_BATCH_SAMPLE", 16)      # the emulated run checks 16 streams against the oracle, not all 1024
    G.test_full_batch_properties_and_sampled_parity(sts, oracle)
test_stats_flag_last_cycle_counters_and_fstats = _pick(G.test_stats_flag_last_cycle_counters_and_fstats, lambda n: True)
test_device_lcg_is_the_reference_generator = _rebind(G.test_device_lcg_is_the_reference_generator, "")
test_device_generators_match_the_reference_tool_and_jump_ahead = _pick(
    G.test_device_generators_match_the_reference_tool_and_jump_ahead, lambda gen: True)
test_chained_generators_are_refused_by_the_device = _rebind(G.test_chained_generators_are_refused_by_the_device, "")
test_cyclic_pattern_histogram = _pick(G.test_cyclic_pattern_histogram, lambda bits: bits in (3, 16, 18))
test_pipelined_batches_complete_in_order = _rebind(G.test_pipelined_batches_complete_in_order, "")
test_assessment_tallies_match_reference_report = _rebind(G.test_assessment_tallies_match_reference_report, "")
test_single_test_parameter_sweep = _pick(G.test_single_test_parameter_sweep, lambda test, prm: prm.get("n", 0) < (1 << 22))
test_partial_test_masks = _rebind(G.test_partial_test_masks, "")
test_error_behaviour = _rebind(G.test_error_behaviour, "")
test_small_streams_against_oracle = _pick(G.test_small_streams_against_oracle, lambda n: True)
test_ascii_input_is_packed_like_the_reference_reads_it = _pick(G.test_ascii_input_is_packed_like_the_reference_reads_it,
                                                               lambda line: True)
test_ascii_input_short_text_and_bad_characters = _rebind(G.test_ascii_input_short_text_and_bad_characters, "")


_SCHEDULE_PROBE = r"""
import hashlib, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from emu_util import emu_binding
from oracle import oracle as O
sts = emu_binding()
h = hashlib.sha256()
for prm in (dict(), dict(n=1000008, linear_complexity_M=1100, serial_m=9, apen_m=7)):
    n = prm.get("n", 1 << 20)
    raw = O.lcg_bytes(40, 2, n)
    with sts.Engine(max_streams=2, **prm) as eng:
        for part in eng.run(raw, 2):
            h.update(part.tobytes())
print(h.hexdigest())
"""


def test_results_do_not_depend_on_the_thread_schedule(sts):
    """a cheap race check: the emulator resumes runnable threads warp 0 / lane 0 first or in a fresh random order on
    every pass (STS_EMU_SCHED; "reverse" exists too); records that differ between the orders would mean a kernel reads
    shared memory another thread writes without a barrier or __syncwarp in between"""
    import subprocess
    import sys
    digests = {}
    for sched in ("forward", "random:3"):
        env = dict(os.environ, STS_EMU_SCHED=sched)
        r = subprocess.run([sys.executable, "-c", _SCHEDULE_PROBE, G.__file__.rsplit("/tests/", 1)[0]], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[sched] = r.stdout.strip()
    assert len(set(digests.values())) == 1, digests


test_repeated_batches_on_one_engine_give_identical_records = _pick(G.test_repeated_batches_on_one_engine_give_identical_records,
                                                                   lambda prm: prm.get("n") in (10 ** 6, 1000008))
test_engine_limits_against_oracle = _pick(G.test_engine_limits_against_oracle, lambda test, prm: test == 14)


def test_apen_sum_with_empty_bins_is_the_sequential_sum(sts, oracle):
    """ApEn's phi = (1/n) sum C log(C/n) in index order (approximateEntropy.c:396-407) on a 7/8-biased stream, whose
    histograms have hundreds of empty bins: the exact parallel summation (rne_sum.cuh) must give the oracle's
    sequential double bit for bit (same libm here).  It used to add one unit in the last place per empty bin that went
    through its fast path - the negated empty bin is -0.0, which was decoded as a number - inside the 1e-13 the GPU
    test allows for libdevice's log, so only the emulator's exact comparison could see it."""
    import numpy as np
    from golden_util import pattern_streams
    n = 1 << 20
    raw = pattern_streams(n, 6).reshape(6, -1)[5].copy()
    for m in (10, 7):
        with sts.Engine(max_streams=1, test_mask=1 << 11, apen_m=m) as eng:
            pv, st, w = eng.run(raw, 1)
            got = st[0, eng.layout.st_off[11]:eng.layout.st_off[11] + 2].copy()
        opv, ost, ow, olay = oracle.run(oracle.make_params(test_mask=1 << 11, apen_m=m), raw, 1)
        want = ost[0, olay.st_off[11]:olay.st_off[11] + 2]
        assert (got == want).all(), (got.view(np.float64), want.copy().view(np.float64))
        assert (pv == opv).all()
test_dft_bluestein_path = _pick(G.test_dft_bluestein_path, lambda n: True)
