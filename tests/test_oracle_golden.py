"""CPU: the oracle (oracle/sts_oracle.c) against the golden vectors produced by the unmodified
reference binary (tests/golden/make_golden.py).  Integer statistics must be identical; p-values must
be bit-identical for the 14 tests that do not depend on libfftw3 and for the DFT too (the golden run
used the same FP64 FFT through the fftw3 shim)."""
import numpy as np
import pytest

from golden_util import check_against_golden, golden_cases, golden_input, golden_params, load_golden

FAST = [c for c in golden_cases() if "config4" not in c]


@pytest.mark.parametrize("case", FAST)
def test_oracle_matches_reference_golden(oracle, case):
    g = load_golden(case)
    prm = oracle.make_params(**golden_params(g))
    raw = golden_input(g, oracle)
    pv, st, w, lay = oracle.run(prm, raw, int(g["streams"]), nthreads=8)
    bad = check_against_golden(g, pv, st, w, lay, pv_rtol=0.0)
    assert not bad, "\n".join(bad)


def test_oracle_matches_reference_golden_config4(oracle):
    """BASELINE config 4 (n = 10^7, LC M=5000, template m=10, Serial/ApEn m=16): one stream, ~40 s"""
    if "lcg_config4_2" not in golden_cases():
        pytest.skip("config-4 golden not generated")
    g = load_golden("lcg_config4_2")
    prm = oracle.make_params(**golden_params(g))
    raw = golden_input(g, oracle)[: int(g["n"]) // 8]
    pv, st, w, lay = oracle.run(prm, raw, 1, nthreads=1)
    sub = {k: (g[k][:1] if g[k].ndim > 0 and k not in ("kind", "cmd") else g[k]) for k in g.files}
    sub["streams"] = np.int64(1)
    bad = check_against_golden(sub, pv, st, w, lay, pv_rtol=0.0)
    assert not bad, "\n".join(bad)


def test_lcg_matches_reference_generator_prefix(oracle):
    """first 64 bytes of `generators 1 1` (captured from the reference tool when the goldens were made)"""
    b = oracle.lcg_bytes(0, 1)
    # known-answer: S_n of stream 0 from the reference's Frequency/stats.txt is -602
    bits = np.unpackbits(b)
    assert int(bits.sum()) * 2 - bits.size == -602


def test_igamc_known_values(oracle):
    from scipy.special import gammaincc
    for a, x in [(3.0, 2.5), (2.5, 1.2), (4.0, 9.0), (32.0, 30.0), (512.0, 500.0), (16384.0, 16400.0), (0.5, 0.3)]:
        assert abs(oracle.igamc(a, x) - gammaincc(a, x)) <= 2e-9 * gammaincc(a, x)
    assert oracle.igamc(3.0, 0.0) == 1.0
    assert oracle.igamc(3.0, 1e6) == 0.0


def test_oracle_assessment_against_reference_report(oracle):
    """oracle_assess (restating X_metrics, frequency.c:631-894) against the reference's finalAnalysisReport.txt
    for 64 LCG streams (tests/golden/make_assess_golden.py); additivity of the tallies over shards."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "assess_lcg_64.npz"))
    streams = 8                                      # the oracle needs ~2.5 s per stream: a prefix, then properties
    raw = oracle.lcg_bytes(0, streams)
    lay = oracle.layout(oracle.make_params())
    pv, st, w, _ = oracle.run(oracle.make_params(), raw, streams, nthreads=os.cpu_count() or 1)
    a = oracle.assess(lay, pv, 10)
    assert (a["sample"] <= streams).all() and (a["freq"].sum(1) == a["sample"]).all()
    assert (a["freq"] <= g["bins"]).all()            # the first 8 of the report's 64 streams
    lo, hi = oracle.assess(lay, pv[:3], 10), oracle.assess(lay, pv[3:], 10)
    for k in ("sample", "toolow", "freq"):
        assert (lo[k] + hi[k] == a[k]).all()
    # known answers: igamc(4.5, chi2 / 2) for a flat and for a one-bin histogram of 10 p-values
    flat = np.tile((np.arange(10) + 0.5) / 10.0, (lay.npv, 1)).T.copy()
    f = oracle.assess(lay, flat, 10)
    e = [c for t in (12, 13) for c in range(lay.pv_off[t], lay.pv_off[t] + lay.pv_cnt[t])]
    assert (f["freq"] == 1).all() and np.allclose(f["uniformity"], 1.0) and (f["passed"] == 10).all()
    assert len(e) == 26


def test_ascii_seek_rule_against_reference_reading(oracle):
    """-F a: the reference seeks to character iteration * n and reads n digits skipping white space
    (utilities.c:1682-1697).  tests/golden/ascii_lines80_3.npz holds its p-values for a text with a line break every
    80 digits; the numpy statement of that rule (golden_util.ascii_seek_rule_streams, which the GPU tests use as the
    expected packing) fed to the oracle must reproduce them bit for bit - a sequential reader would not."""
    import os
    from golden_util import GOLDEN_DIR, ascii_lines, ascii_seek_rule_streams
    g = np.load(os.path.join(GOLDEN_DIR, "ascii_lines80_3.npz"))
    n, it = int(g["n"]), int(g["iterations"])
    bits = np.unpackbits(oracle.lcg_bytes(0, int(g["source_streams"])))
    text = ascii_lines(bits, int(g["line"]))
    packed, complete = ascii_seek_rule_streams(text, n, it)
    assert complete == it
    assert not (packed[1] == np.packbits(bits[n:2 * n])).all()        # the overlap: stream 1 is NOT bits [n, 2n)
    mask = (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 9) | (1 << 12) | (1 << 13)     # the cheap tests
    pv, st, w, lay = oracle.run(oracle.make_params(test_mask=mask), packed.reshape(-1), it, nthreads=it)
    for t in range(1, 16):
        if (mask >> t) & 1:
            mine = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
            assert (mine == g["pv_%d" % t]).all(), "test %d" % t
