"""CPU: how firmly the DFT test's N_1 is pinned although libfftw3 (the reference's FFT, un-vendored and absent
here - SURVEY 8c "parity unpinned") cannot be run.

The golden N_1 values come from the unmodified reference linked against oracle/fft_f64.c through the fftw3.h shim.
Here an INDEPENDENT FP64 FFT library (numpy's pocketfft) recomputes the moduli of the same streams as
discreteFourierTransform.c:392-411 defines them (m_i = |X_i| for i in [0, n/2), counted when m_i < T), and the
test asserts, per golden stream,
  1. pocketfft's N_1 equals the golden N_1,
  2. the two FFT implementations disagree on a modulus by less than 1e-12 * T (measured: a few 1e-15 * T), and
  3. the modulus closest to T is more than 1e-9 * T away from it (the band in which the north star lets N_1
     differ) and more than 1000 x the measured disagreement of the two libraries -
so any correct FP64 FFT, FFTW included, yields these N_1, and the GPU's squared-threshold comparison
(k_dft_fast.cu) cannot flip a bin on these inputs either.  (With 5 * 10^5 .. 5 * 10^6 moduli spread around T the
closest one is typically 1e-6 .. 1e-8 of T away: 8.6e-9 on the first n = 10^7 stream.)"""
import ctypes as C

import numpy as np
import pytest

from golden_util import golden_cases, golden_input, load_golden


def oracle_moduli(O, x):
    """|r2c(x)|[0 : n/2] with oracle/fft_f64.c, the FFT behind the golden values"""
    lib = O.lib()
    lib.fft_plan_create.restype = C.c_void_p
    lib.fft_plan_create.argtypes = [C.c_long]
    lib.fft_plan_destroy.argtypes = [C.c_void_p]
    lib.fft_r2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    n = x.size
    out = np.zeros((n // 2 + 1, 2), dtype=np.float64)
    plan = lib.fft_plan_create(n)
    assert plan
    lib.fft_r2c(plan, x.ctypes.data, out.ctypes.data)
    lib.fft_plan_destroy(plan)
    return np.hypot(out[: n // 2, 0], out[: n // 2, 1])


@pytest.mark.parametrize("case", golden_cases())
def test_peak_count_is_the_same_for_an_independent_fft_and_clear_of_the_threshold(oracle, case):
    g = load_golden(case)
    if not (int(g["enabled_mask"]) >> 7) & 1:
        pytest.skip("DFT disabled in this case")
    n, streams = int(g["n"]), int(g["streams"])
    raw = golden_input(g, oracle).reshape(streams, n // 8)
    T = np.sqrt(np.log(20.0) * n)                       # discreteFourierTransform.c:142, log(1 / 0.05) = log(20)
    for s in range(min(streams, 4 if n <= (1 << 21) else 1)):
        x = np.unpackbits(raw[s]).astype(np.float64) * 2.0 - 1.0        # utilities.c:1862-1872 MSB first; X = 2 eps - 1
        m = np.abs(np.fft.rfft(x))[: n // 2]            # i = 0 .. n/2 - 1: DC included, Nyquist excluded (quirk A4)
        n1 = int((m < T).sum())
        assert n1 == int(g["dft_N1"][s]), "stream %d: pocketfft N_1 %d, reference %d" % (s, n1, int(g["dft_N1"][s]))
        disagreement = float(np.max(np.abs(m - oracle_moduli(oracle, x))) / T)
        assert disagreement < 1e-12, disagreement
        margin = float(np.min(np.abs(m - T)) / T)
        assert margin > 1e-9 and margin > 1000.0 * disagreement, \
            "stream %d: a modulus lies within %.3g * T of T (FFT disagreement %.3g * T)" % (s, margin, disagreement)
