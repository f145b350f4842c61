"""GPU (-m gpu): the CUDA engine, called through the C ABI (libstsgpu.so via ctypes), against
  (1) the golden vectors produced by the unmodified reference binary (tests/golden/*.npz),
  (2) the CPU oracle on the same seeded inputs, and
  (3) size-independent properties at BASELINE's full batch size (1024 x 2^20-bit streams).

Bars: integer statistics, template counts, histograms bit-exact; p-values within PV_RTOL = 1e-9 relative
(the north-star tolerance) - in practice ~1e-15, the residue of CUDA libdevice vs glibc erf/erfc/exp/log.
CumulativeSums p-values are formed as 1 - sum1 + sum2 (cusum.c:369), so a p-value near 0 carries the
absolute rounding noise of sums of magnitude 1; for that test only an absolute floor PV_ATOL_CUSUM
applies on top of the relative bar."""
import os

import numpy as np
import pytest

from golden_util import check_against_golden, golden_cases, golden_input, golden_params, load_golden, rel_err

pytestmark = pytest.mark.gpu

PV_RTOL = 1e-9
PV_ATOL_CUSUM = 1e-13


@pytest.fixture(scope="module")
def sts():
    import sts_b200
    sts_b200.load_library()          # raises if the CUDA extension is missing: no fallback
    return sts_b200


def half_is_5_smooth(n):
    h = n // 2
    for p in (2, 3, 5):
        while h % p == 0:
            h //= p
    return h == 1


def engine_for(sts, g, streams, mask=None):
    prm = golden_params(g)
    n = prm["n"]
    m = sts.ALL_TESTS if mask is None else mask
    if not sts.dft_supported(n) or not half_is_5_smooth(n):
        # beyond the engine's FFT lengths it refuses the DFT loudly; the Bluestein lengths have their own
        # test at the end of this file (test_dft_bluestein_path), so that the four-step kernels are judged first
        m &= ~(1 << 7)
    return sts.Engine(max_streams=streams, test_mask=m, **prm), m


def compare_with_oracle(sts, oracle, eng, mask, prm, raw, streams, pv, st, w):
    lay = eng.layout
    opv, ost, ow, olay = oracle.run(oracle.make_params(test_mask=mask, **prm), raw, streams,
                                    nthreads=os.cpu_count() or 1)
    assert lay.enabled_mask == olay.enabled_mask
    assert (lay.npv, lay.nst, lay.nw) == (olay.npv, olay.nst, olay.nw)
    assert list(lay.pv_off) == list(olay.pv_off) and list(lay.st_off) == list(olay.st_off)
    assert (w == ow).all(), "W[template][block] differs"
    fp = {lay.st_off[10]} if (lay.enabled_mask >> 10) & 1 else set()
    if (lay.enabled_mask >> 11) & 1:
        fp |= {lay.st_off[11], lay.st_off[11] + 1}
    cols = [c for c in range(lay.nst) if c not in fp]
    bad = np.argwhere(st[:, cols] != ost[:, cols])
    assert bad.size == 0, "integer statistics differ at (stream, slot) %s" % bad[:8].tolist()
    for c in fp:                     # double bit patterns: Universal sum (bit-exact), ApEn phi (FP)
        a = st[:, c].copy().view(np.float64)
        b = ost[:, c].copy().view(np.float64)
        if c == lay.st_off[10] and (lay.enabled_mask >> 10) & 1:
            assert (a == b).all(), "Universal sum must be bit-identical (sequential order, host log table)"
        else:
            assert rel_err(a, b).max() <= 1e-13
    for t in range(1, 16):
        if not (lay.enabled_mask >> t) & 1:
            continue
        a = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
        b = opv[:, olay.pv_off[t]:olay.pv_off[t] + olay.pv_cnt[t]]
        r = rel_err(a, b)
        if t == 3:
            r[np.abs(a - b) <= PV_ATOL_CUSUM] = 0.0
        assert r.max() <= PV_RTOL, "test %d (%s): p-value rel err %.3g" % (t, sts.TEST_NAMES[t], r.max())


@pytest.mark.parametrize("case", [c for c in golden_cases() if "config4" not in c])
def test_engine_matches_reference_golden_and_oracle(sts, oracle, case):
    g = load_golden(case)
    streams = int(g["streams"])
    raw = golden_input(g, oracle)
    eng, mask = engine_for(sts, g, streams)
    with eng:
        pv, st, w = eng.run(raw, streams)
        lay = eng.layout
        gg = {k: g[k] for k in g.files}
        gg["enabled_mask"] = np.int64(int(g["enabled_mask"]) & mask)
        bad = check_against_golden(gg, pv, st, w, lay, pv_rtol=PV_RTOL)
        # cusum absolute floor (see module docstring)
        bad = [b for b in bad if not b.startswith("test 3:")] + \
              [b for b in bad if b.startswith("test 3:") and
               np.abs(pv[:, lay.pv_off[3]:lay.pv_off[3] + 2] - g["pv_3"]).max() > PV_ATOL_CUSUM]
        assert not bad, "\n".join(bad)
        compare_with_oracle(sts, oracle, eng, mask, golden_params(g), raw, streams, pv, st, w)


def test_engine_config4_parameters(sts, oracle):
    """BASELINE config 4: n = 10^7, LC M = 5000 (the `1 << M` class rule), template m = 10, Serial/ApEn
    m = 16 (two histogram slices of 2^16 counters each, 3 * 2^16 ApEn bins summed in index order)."""
    if "lcg_config4_2" not in golden_cases():
        pytest.skip("config-4 golden not generated")
    g = load_golden("lcg_config4_2")
    raw = golden_input(g, oracle)
    eng, mask = engine_for(sts, g, 2)
    with eng:
        pv, st, w = eng.run(raw, 2)
        lay = eng.layout
        assert not (lay.enabled_mask >> 2) & 1      # BlockFrequency self-disables (blockFrequency.c:119)
        gg = {k: g[k] for k in g.files}
        gg["enabled_mask"] = np.int64(int(g["enabled_mask"]) & mask)
        bad = check_against_golden(gg, pv, st, w, lay, pv_rtol=PV_RTOL)
        assert not bad, "\n".join(bad)


@pytest.mark.parametrize("n", [1 << 20, 1000008])
def test_stats_flag_last_cycle_counters_and_fstats(sts, oracle, n):
    """STSGPU_FLAG_STATS (-s): the visit counts of the LAST excursion cycle (randomExcursions.c:386, :860-866)
    bit-exact against the oracle, fstats consistent with the integers and p-values of the same record, and every other
    column unchanged by the flag.  (That the -s files made from them equal the reference's is tests/test_integration.py.)"""
    B = 6
    raw = oracle.lcg_bytes(3, B, n)
    mask = sts.ALL_TESTS & ~(1 << 7) if not half_is_5_smooth(n) else sts.ALL_TESTS
    with sts.Engine(max_streams=B, n=n, test_mask=mask, flags=sts.FLAG_STATS) as eng:
        pv, st, w = eng.run(raw, B)
        fs = eng.results_stats()
        lay = eng.layout
    with sts.Engine(max_streams=B, n=n, test_mask=mask) as eng:
        pv0, st0, w0 = eng.run(raw, B)
        assert eng.results_stats() is None
    o12 = lay.st_off[12]
    assert (pv == pv0).all() and (w == w0).all()
    keep = [c for c in range(lay.nst) if not (o12 + 49 <= c < o12 + 57)]
    assert (st[:, keep] == st0[:, keep]).all() and (st0[:, o12 + 49:o12 + 57] == 0).all()
    opv, ost, ow, olay = oracle.run(oracle.make_params(n=n, test_mask=mask, flags=1), raw, B, nthreads=os.cpu_count() or 1)
    assert (st[:, o12 + 49:o12 + 57] == ost[:, o12 + 49:o12 + 57]).all()
    possible = st[:, o12] >= 500
    assert possible.any() and (st[possible, o12 + 49:o12 + 57].sum(axis=1) > 0).all()
    assert fs.shape == (B, lay.nfs) and lay.nfs == sum(lay.fs_cnt) == 24 + lay.ntemplates + 2 * ((lay.enabled_mask >> 7) & 1)
    # Serial: del1 = psim0 - psim1, del2 = psim0 - 2 psim1 + psim2 (serial.c:217-218)
    f14 = fs[:, lay.fs_off[14]:lay.fs_off[14] + 5]
    assert (f14[:, 3] == f14[:, 0] - f14[:, 1]).all() and (f14[:, 4] == f14[:, 0] - 2.0 * f14[:, 1] + f14[:, 2]).all()
    # Runs: pi = ones / n (runs.c:219); DFT: N_0 = 0.95 n / 2; Rank: p = exp(-chi2 / 2) (rank.c:346)
    assert (fs[:, lay.fs_off[4]] == st[:, lay.st_off[4]] / float(n)).all()
    if (lay.enabled_mask >> 7) & 1:
        assert (fs[:, lay.fs_off[7]] == 0.95 * n / 2.0).all()
    assert rel_err(np.exp(-fs[:, lay.fs_off[6]] / 2.0), pv[:, lay.pv_off[6]]).max() <= 1e-12
    # ApEn: ApEn = phi(m) - phi(m+1) (approximateEntropy.c:222)
    phi = st[:, lay.st_off[11]:lay.st_off[11] + 2].copy().view(np.float64)
    assert (fs[:, lay.fs_off[11]] == phi[:, 0] - phi[:, 1]).all()
    # non-overlapping: chi2 = sum ((W - mu) / sigma)^2 per template (nonOverlappingTemplateMatchings.c:561-565)
    m, M = 9, n // 8
    mu = (M - m + 1) / float(1 << m)
    sig = np.sqrt(M * (1.0 / (1 << m) - (2.0 * m - 1.0) / (1 << 2 * m)))
    chi = (((w.reshape(B, -1, 8).astype(np.float64) - mu) / sig) ** 2).sum(axis=2)
    assert rel_err(fs[:, lay.fs_off[8]:lay.fs_off[8] + lay.ntemplates], chi).max() <= 1e-12


def test_device_lcg_is_the_reference_generator(sts, oracle):
    with sts.Engine(max_streams=8) as eng:
        eng.generate_lcg(0, 8)
        assert (eng.download_input(8) == oracle.lcg_bytes(0, 8)).all()
        eng.generate_lcg(1000, 4)                    # skip-ahead: golden case lcg_default_off1000_4
        assert (eng.download_input(4) == oracle.lcg_bytes(1000, 4)).all()
    with sts.Engine(max_streams=2, n=1000008, test_mask=sts.ALL_TESTS & ~(1 << 7)) as eng:
        eng.generate_lcg(1, 2)                       # ragged stream length, mid-file start
        assert (eng.download_input(2) == oracle.lcg_bytes(1, 2, 1000008)).all()


@pytest.mark.parametrize("gen", [2, 4, 5])
def test_device_generators_match_the_reference_tool_and_jump_ahead(sts, gen):
    """generators 2, 4, 5 of tools/generators.c made on the device (csrc/k_gen.cu): the first streams against the
    vectors of the unmodified tool, then streams far into the tool's output and ragged lengths against the oracle
    (oracle/generators.py, itself pinned to the same vectors) - every stream and every segment of a stream starts
    from its own jump, so a stream made alone equals the same stream made as part of a longer run"""
    import hashlib
    import json
    from oracle import generators as OG
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generators.json")))[str(gen)]
    k = len(gold["sha256"])
    with sts.Engine(max_streams=4) as eng:
        eng.generate(gen, 0, k)
        got = eng.download_input(k).reshape(k, -1)
        assert [hashlib.sha256(got[i].tobytes()).hexdigest() for i in range(k)] == gold["sha256"]
        eng.generate(gen, k - 1, 1)                  # the last of them again, from its own jump
        assert hashlib.sha256(eng.download_input(1).tobytes()).hexdigest() == gold["sha256"][k - 1]
    for n, first, cnt in ((1024 * 40, 37, 3), (1000 * 8, 1, 2), (512 * 3 + 8, 5, 4), (4096 + 24, 0, 3)):
        want = OG.generate(gen, n, first + cnt)[first:]
        with sts.Engine(max_streams=4, n=n, test_mask=1 << 1) as eng:
            eng.generate(gen, first, cnt)
            got = eng.download_input(cnt).reshape(cnt, -1)
        assert [got[i].tobytes() for i in range(cnt)] == want, (gen, n, first)


def test_chained_generators_are_refused_by_the_device(sts):
    with sts.Engine(max_streams=1) as eng:
        for gen in (3, 6, 7, 8, 9):
            with pytest.raises(sts.EngineError, match="host"):
                eng.generate(gen, 0, 1)
        eng.generate(1, 0, 1)                        # generator 1 is the LCG call


@pytest.mark.parametrize("bits", [3, 9, 11, 16, 18])
def test_cyclic_pattern_histogram(sts, oracle, bits):
    raw = oracle.lcg_bytes(3, 1)
    with sts.Engine(max_streams=1) as eng:
        eng.upload(raw, 1)
        h = eng.debug_pattern_histogram(0, bits)
    assert (h == oracle.pattern_histogram(raw, 1 << 20, bits)).all()
    assert int(h.sum()) == 1 << 20


FULL_BATCH_SAMPLE = 0


def test_full_batch_properties_and_sampled_parity(sts, oracle):
    """BASELINE config 2: 1024 streams.  Properties that hold for any input + oracle parity on a sample."""
    B = 1024
    n = 1 << 20
    with sts.Engine(max_streams=B) as eng:
        eng.generate_lcg(0, B)
        eng.submit_resident(B, 0)
        eng.wait()
        pv, st, w = eng.results()
        lay = eng.layout
        o = lay.st_off
        ones = st[:, o[4]]
        assert (st[:, o[1]] == 2 * ones - n).all()                         # S_n = #1 - #0
        assert (st[:, o[2]] == 64).all() and (st[:, o[2] + 1:o[2] + 65].sum(1) == ones).all()
        assert (st[:, o[3]] == st[:, o[1]]).all()                          # cusum final S
        assert (st[:, o[3] + 1] >= np.maximum(st[:, o[3]], 0)).all() and (st[:, o[3] + 2] <= 0).all()
        assert (st[:, o[5]:o[5] + 7].sum(1) == 104).all()                  # longest-run blocks
        assert (st[:, o[6]:o[6] + 3].sum(1) == 1024).all()                 # matrices
        assert ((st[:, o[7]] > 0) & (st[:, o[7]] <= n // 2)).all()
        assert (st[:, o[9]:o[9] + 6].sum(1) == 1016).all()                 # overlapping blocks
        J = st[:, o[12]]
        v = st[:, o[12] + 1:o[12] + 49].reshape(B, 6, 8)
        assert (v.sum(1) == J[:, None]).all()                              # every cycle lands in one class per state
        assert (st[:, o[13]] == J).all()
        assert (st[:, o[14]] >= n * n // 65536).all()                      # sum v^2 >= n^2 / bins
        assert (st[:, o[15]:o[15] + 7].sum(1) == 2097).all()               # LC blocks
        # template counts: every block's windows are counted at most once per template
        assert (w.reshape(B, 148, 8).sum(1) <= 131072 - 8).all()
        # p-values are in [0, 1] or NON_P_VALUE
        okp = ((pv >= 0) & (pv <= 1.0 + 1e-12)) | (pv == sts.NON_P_VALUE)
        assert okp.all()
        # determinism / idempotence
        eng.submit_resident(B, 0)
        eng.wait()
        pv2, st2, w2 = eng.results()
        assert (pv2 == pv).all() and (st2 == st).all() and (w2 == w).all()
        raw = eng.download_input(B)
    # batch-split invariance: streams 512..575 alone give the same records
    with sts.Engine(max_streams=64) as eng:
        p64, s64, w64 = eng.run(raw[512 * n // 8:576 * n // 8].copy(), 64, 512)
        assert (p64 == pv[512:576]).all() and (s64 == st[512:576]).all() and (w64 == w[512:576]).all()
    # oracle parity on EVERY stream of the batch (BASELINE config 2: "1024 x 2^20 ... bit-exact integer stats + p-value
    # check vs reference"): about 2.6 s of oracle per stream and host thread, ~3 min on the 16 cores of a GPU box.
    # FULL_BATCH_SAMPLE = k checks k streams spread over the batch instead (the CPU emulator's run of this body).
    k = FULL_BATCH_SAMPLE
    idx = np.linspace(0, B - 1, k).astype(int) if 0 < k < B else np.arange(B)
    sub = raw if len(idx) == B else np.concatenate([raw[i * n // 8:(i + 1) * n // 8] for i in idx])
    opv, ost, ow, olay = oracle.run(oracle.make_params(), sub, len(idx), nthreads=os.cpu_count() or 1)
    assert (ow == w[idx]).all(), "W[template][block] differs in streams %s" % idx[(ow != w[idx]).any(axis=1)][:8].tolist()
    fp = [lay.st_off[11], lay.st_off[11] + 1]
    cols = [c for c in range(lay.nst) if c not in fp]
    bad = np.argwhere(ost[:, cols] != st[idx][:, cols])
    assert bad.size == 0, "integer statistics differ at (stream, slot) %s" % bad[:8].tolist()
    phi_g = st[idx][:, fp].copy().view(np.float64)
    phi_o = ost[:, fp].copy().view(np.float64)
    assert rel_err(phi_g, phi_o).max() <= 1e-13
    r = rel_err(pv[idx], opv)
    c3 = slice(lay.pv_off[3], lay.pv_off[3] + 2)
    r[:, c3][np.abs(pv[idx][:, c3] - opv[:, c3]) <= PV_ATOL_CUSUM] = 0.0
    assert r.max() <= PV_RTOL, "p-values differ: stream %d slot %d" % np.unravel_index(np.argmax(r), r.shape)
    assert (pv[idx] == sts.NON_P_VALUE).sum() == (opv == sts.NON_P_VALUE).sum() > 0


def test_pipelined_batches_complete_in_order(sts, oracle):
    """pipeline_depth = 2: two batches in flight give the records of the same batches run one at a time"""
    n = 1 << 20
    raw = oracle.lcg_bytes(20, 6)
    a, b, c = raw[:2 * n // 8].copy(), raw[2 * n // 8:4 * n // 8].copy(), raw[4 * n // 8:].copy()
    with sts.Engine(max_streams=2) as one:
        ref = [one.run(x, 2, 100 + 2 * i) for i, x in enumerate((a, b, c))]
    with sts.Engine(max_streams=2, pipeline_depth=2) as eng:
        eng.submit(a, 2, 100)
        eng.submit(b, 2, 102)
        assert eng.pending() == 2
        with pytest.raises(sts.EngineError):
            eng.submit(c, 2, 104)                            # third batch while two are in flight
        got = []
        eng.wait()
        got.append(eng.results())
        eng.submit(c, 2, 104)                                # slot of batch a is free again
        eng.wait()
        got.append(eng.results())
        eng.wait()
        got.append(eng.results())
        assert eng.pending() == 0
        with pytest.raises(sts.EngineError):
            eng.wait()
    for (pv, st, w), (rpv, rst, rw) in zip(got, ref):
        assert (pv == rpv).all() and (st == rst).all() and (w == rw).all()


def test_assessment_tallies_match_reference_report(sts, oracle):
    """stsgpu_assess_*: per-column uniformity bins, proportion and uniformity p-value of 64 LCG streams, tallied
    on the device batch by batch (two batches in flight), against the reference's own finalAnalysisReport.txt
    (sts -O: 10 bins; tests/golden/make_assess_golden.py) and against the oracle's restatement of X_metrics."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "assess_lcg_64.npz"))
    streams, B, n = int(g["streams"]), 16, 1 << 20
    raw = oracle.lcg_bytes(0, streams)
    pvs = []
    with sts.Engine(max_streams=B, pipeline_depth=2) as eng:
        eng.assess_begin(10)
        pending = 0
        for b in range(streams // B):
            eng.submit(raw[b * B * n // 8:(b + 1) * B * n // 8].copy(), B, b * B)
            pending += 1
            if pending == 2:
                eng.wait()
                pvs.append(eng.results()[0])
                pending -= 1
        while pending:
            eng.wait()
            pvs.append(eng.results()[0])
            pending -= 1
        a = eng.assess_finish()
        lay = eng.layout
        # a different bin count restarts the tallies: sqrt(iterations) bins, parse_args.c:791-793
        eng.assess_begin(8, 0.01, 0.0001)
        eng.run(raw[:B * n // 8].copy(), B, 0)
        a8 = eng.assess_finish()
    pv = np.concatenate(pvs)
    assert (a["freq"] == g["bins"]).all() and (a["sample"] == g["sample"]).all() and (a["passed"] == g["passed"]).all()
    assert np.abs(a["uniformity"] - g["uniformity"]).max() <= 5.01e-7          # the report prints %8.6f
    assert ((a["flags"] & 1) == 0).sum() == g["uniformity_failed"].sum()
    assert ((a["flags"] & 2) == 0).sum() == g["proportion_failed"].sum()
    o = oracle.assess(oracle.layout(oracle.make_params()), pv, 10)
    for k in ("sample", "toolow", "passed", "freq", "flags"):
        assert (a[k] == o[k]).all(), k
    for k in ("tmin", "tmax", "chi2", "uniformity"):
        assert rel_err(a[k], o[k]).max() <= PV_RTOL, k
    o8 = oracle.assess(oracle.layout(oracle.make_params()), pv[:B], 8)
    assert (a8["freq"] == o8["freq"]).all() and rel_err(a8["uniformity"], o8["uniformity"]).max() <= PV_RTOL
    assert lay.npv == 188


@pytest.mark.parametrize("test,prm", [
    (15, dict(n=1 << 20, linear_complexity_M=1100)),     # LinearComplexity: warp-per-block kernel, 2 words per lane
    (15, dict(n=1 << 20, linear_complexity_M=2048)),     # ... 3 words per lane (W = 65)
    (15, dict(n=1 << 20, linear_complexity_M=3100)),     # ... 4
    (15, dict(n=1 << 20, linear_complexity_M=4100)),     # ... 5
    (15, dict(n=1 << 20, linear_complexity_M=1022)),     # last M of the per-thread register kernel (W = 32)
    (15, dict(n=1 << 20, linear_complexity_M=511)),      # last M of the windowed kernel (W = 16)
    (10, dict(n=2200000)),                               # Universal L = 8
    (10, dict(n=904960)),                                # Universal L = 7 at its smallest n
    (14, dict(n=1 << 18, serial_m=3)),                   # Serial: smallest block length
    (14, dict(n=1 << 18, serial_m=12)),
    (11, dict(n=1 << 18, apen_m=1)),                     # ApEn: H = 2
    (11, dict(n=1 << 20, apen_m=14)),                    # ApEn: H = 15
    (14, dict(n=1 << 21, serial_m=18)),                  # Serial: sliced 32-bit histogram, 8 slices
    (14, dict(n=(1 << 20) + 40, serial_m=16)),           # Serial: the 17-bit-window kernel with a partial last group (n % 32 = 8)
    (14, dict(n=1 << 20, serial_m=15)),                  # Serial: H = 15, the 16-bit-counter kernel
    (8, dict(n=1 << 18, non_overlapping_m=8)),
    (8, dict(n=1 << 18, non_overlapping_m=12)),              # wide-window stride 2
    (8, dict(n=(1 << 18) + 8 * 13, non_overlapping_m=10)),   # stride 4 with a partial last group of window starts
    (8, dict(n=1 << 18, non_overlapping_m=13)),              # stride 1 (2^13 bins already)
    (8, dict(n=(1 << 19) - 8 * 5, non_overlapping_m=15)),    # largest m
    (2, dict(n=1 << 20, block_frequency_M=11000)),       # blocks that are not word aligned
    (7, dict(n=1 << 14)),                                # DFT: generic path, small power of two
    (7, dict(n=5 * (1 << 16))),                          # DFT: one radix-5 pass
    (7, dict(n=1 << 22)),                                # DFT: generic path above the fast kernel's length
    (7, dict(n=1 << 25)),                                # DFT: 4096 x 4096, the largest split
    (7, dict(n=3 * 5 * (1 << 15))),                      # DFT: radices 3 and 5 together
])
def test_single_test_parameter_sweep(sts, oracle, test, prm):
    """kernel variants the default parameters never reach, one test at a time, bit-exact / 1e-9 against the oracle"""
    base = dict(n=1 << 20, block_frequency_M=16384, non_overlapping_m=9, overlapping_m=9, apen_m=10, serial_m=16,
                linear_complexity_M=500)
    base.update(prm)
    if test == 2 and base["n"] // base["block_frequency_M"] >= 100:
        pytest.skip("BlockFrequency disables itself")
    streams, mask = 2, 1 << test
    raw = oracle.lcg_bytes(3, streams, base["n"])
    with sts.Engine(max_streams=streams, test_mask=mask, **base) as eng:
        assert eng.layout.enabled_mask == mask == oracle.layout(oracle.make_params(test_mask=mask, **base)).enabled_mask
        pv, st, w = eng.run(raw, streams)
        compare_with_oracle(sts, oracle, eng, mask, base, raw, streams, pv, st, w)


def test_partial_test_masks(sts, oracle):
    """state->testVector subsets (the reference's -t flag) change the record layout, not the values"""
    raw = oracle.lcg_bytes(5, 2)
    with sts.Engine(max_streams=2) as full:
        pv_all, st_all, w_all = full.run(raw, 2)
        lay_all = full.layout
        for mask in (1 << 1, (1 << 8) | (1 << 15), (1 << 7) | (1 << 11) | (1 << 14), (1 << 12), (1 << 13) | (1 << 3)):
            with sts.Engine(max_streams=2, test_mask=mask) as eng:
                pv, st, w = eng.run(raw, 2)
                lay = eng.layout
                assert lay.enabled_mask == mask
                for t in range(1, 16):
                    if (mask >> t) & 1:
                        a = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
                        b = pv_all[:, lay_all.pv_off[t]:lay_all.pv_off[t] + lay_all.pv_cnt[t]]
                        assert (a == b).all(), "test %d differs when run alone" % t


def test_error_behaviour(sts):
    with pytest.raises(sts.EngineError):
        sts.Engine(max_streams=1, n=1001)                    # n % 8 != 0  (parse_args.c:939)
    with pytest.raises(sts.EngineError):
        sts.Engine(max_streams=1, n=(1 << 25) + 8)           # DFT beyond 2^25 bits with n/2 not 5-smooth: refused loudly
    with sts.Engine(max_streams=2) as eng:
        with pytest.raises(sts.EngineError):
            eng.wait()                                       # wait without submit
        with pytest.raises(sts.EngineError):
            eng.submit_resident(3, 0)                        # over capacity
        eng.generate_lcg(0, 2)
        eng.submit_resident(2, 0)
        with pytest.raises(sts.EngineError):
            eng.submit_resident(2, 0)                        # batch still in flight
        eng.wait()


@pytest.mark.parametrize("n", [2048, 65536, 1 << 17, 400000, 6 * (1 << 17), 10 ** 6])
def test_small_streams_against_oracle(sts, oracle, n):
    """short streams: most tests self-disable exactly like the X_init checks; the rest must still agree"""
    streams = 3
    raw = oracle.lcg_bytes(2, streams, n)
    mask = sts.ALL_TESTS if sts.dft_supported(n) else sts.ALL_TESTS & ~(1 << 7)
    prm = dict(n=n, block_frequency_M=16384, non_overlapping_m=9, overlapping_m=9, apen_m=10, serial_m=16,
               linear_complexity_M=500)
    with sts.Engine(max_streams=streams, test_mask=mask, **prm) as eng:
        assert eng.layout.enabled_mask == oracle.layout(oracle.make_params(test_mask=mask, **prm)).enabled_mask
        pv, st, w = eng.run(raw, streams)
        compare_with_oracle(sts, oracle, eng, mask, prm, raw, streams, pv, st, w)


def ascii_text(bits, line):
    """'0'/'1' characters with a line break after every `line` digits (none if line == 0)"""
    ch = (bits + ord("0")).astype(np.uint8)
    if not line:
        return ch
    rows = -(-ch.size // line)
    pad = np.full(rows * line, ord(" "), dtype=np.uint8)
    pad[:ch.size] = ch
    out = np.concatenate([pad.reshape(rows, line), np.full((rows, 1), ord("\n"), dtype=np.uint8)], axis=1).reshape(-1)
    return out


def seek_rule_streams(text, n, streams):
    """the reference's reading of an ASCII file, pinned to the reference binary by
    tests/test_oracle_golden.py::test_ascii_seek_rule_against_reference_reading"""
    from golden_util import ascii_seek_rule_streams
    return ascii_seek_rule_streams(text, n, streams)


@pytest.mark.parametrize("line", [0, 64, 61, 1])
def test_ascii_input_is_packed_like_the_reference_reads_it(sts, oracle, line):
    """-F a on the device (stsgpu_submit_ascii): packed rows bit-exact against the reference's seek rule, records equal
    to the raw path on those rows; line breaks every 64 / 61 / 1 digits, or none"""
    n, streams = 1 << 16, 5
    mask = (1 << 1) | (1 << 3) | (1 << 4) | (1 << 12) | (1 << 13) | (1 << 14)
    bits = np.unpackbits(oracle.lcg_bytes(0, 1))[:n * (streams + 1) + 999]
    text = ascii_text(bits, line)[:(streams + 1) * n]
    want, complete = seek_rule_streams(text, n, streams)
    assert complete == streams
    eng = sts.Engine(max_streams=streams, n=n, test_mask=mask)
    eng.submit_ascii(text, streams, 0)
    eng.wait()
    assert eng.ascii_status() == (streams, -1)
    pv_a, st_a, _ = eng.results()
    got = eng.download_input(streams).reshape(streams, n // 8)
    assert (got == want).all(), "packed rows differ from the reference's reading of the text"
    eng.submit(want.reshape(-1), streams, 0)
    eng.wait()
    pv_r, st_r, _ = eng.results()
    assert (st_a == st_r).all() and (pv_a == pv_r).all()
    if line == 0:                               # without white space the streams are the file's bits in sequence
        assert (want.reshape(-1) == np.packbits(bits[:n * streams])).all()
    eng.close()


def test_ascii_input_short_text_and_bad_characters(sts, oracle):
    n, streams = 1 << 16, 4
    bits = np.unpackbits(oracle.lcg_bytes(0, 1))[:n * 3 + 100]
    text = ascii_text(bits, 64)
    _, complete = seek_rule_streams(text, n, streams)
    assert 0 < complete < streams
    eng = sts.Engine(max_streams=streams, n=n, test_mask=1 << 1)
    eng.submit_ascii(text, streams, 0)
    eng.wait()
    assert eng.ascii_status() == (complete, -1)
    assert eng.results()[0].shape[0] == complete          # only the complete leading streams are reported
    bad = text.copy()
    bad[70000] = ord("x")
    bad[100] = ord("2")                          # the reference would store a 2 in epsilon; here it is an error
    eng.submit_ascii(bad, 1, 0)
    eng.wait()
    assert eng.ascii_status() == (1, 100)
    late = text.copy()
    late[n + n // 64 + 5] = ord("x")             # after stream 0's last digit: the reference never scans it
    eng.submit_ascii(late, 1, 0)
    eng.wait()
    assert eng.ascii_status() == (1, -1)
    eng.close()


def test_contexts_are_independent_and_release_their_memory(sts, oracle):
    """two contexts with different parameters side by side give the results each gives alone, and creating /
    destroying contexts (DFT work buffers, ASCII staging, tallies included) returns the device memory"""
    import torch
    raw = oracle.lcg_bytes(0, 2)
    with sts.Engine(max_streams=2) as a, sts.Engine(max_streams=2, n=1 << 19, linear_complexity_M=1500,
                                                     test_mask=(1 << 7) | (1 << 15) | (1 << 1)) as b:
        a.submit(raw, 2, 0)
        b.submit(raw, 2, 0)                       # reads the first 2 x 2^19 bits
        b.wait()
        a.wait()
        pa, sa, wa = a.results()
        pb, sb, _ = b.results()
    with sts.Engine(max_streams=2) as a:
        pa1, sa1, wa1 = a.run(raw, 2, 0)
    assert (pa == pa1).all() and (sa == sa1).all() and (wa == wa1).all()
    with sts.Engine(max_streams=2, n=1 << 19, linear_complexity_M=1500, test_mask=(1 << 7) | (1 << 15) | (1 << 1)) as b:
        pb1, sb1, _ = b.run(raw, 2, 0)
    assert (pb == pb1).all() and (sb == sb1).all()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for k in range(6):
        with sts.Engine(max_streams=64, pipeline_depth=3) as e:
            e.assess_begin(8)
            e.generate_lcg(0, 4)
            e.submit_resident(4, 0)
            e.wait()
            e.submit_ascii(np.full(2 << 20, ord("1"), dtype=np.uint8), 1, 0)
            e.wait()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (64 << 20), "device memory not returned: %d MiB" % ((free0 - free1) >> 20)


@pytest.mark.parametrize("prm", [
    dict(),                                                              # n = 2^20: the specialised FFT kernels
    dict(n=10 ** 6),                                                     # n = 10^6: 200-thread CTAs (partial last warp)
    dict(n=1000008, linear_complexity_M=1100, serial_m=9, apen_m=7),     # ragged words, warp-per-block LC
    dict(n=1 << 17),                                                     # short streams: most tests disabled
])
def test_repeated_batches_on_one_engine_give_identical_records(sts, oracle, prm):
    """a context is reused for every batch of a run: the records of a batch must not depend on what the engine ran
    before (stale shared memory, counters that are not reset, work buffers of the previous batch).  Three batches -
    A, B, A again - on a pipelined engine, the second A must equal the first bit for bit, and A must equal a fresh
    engine's answer.  (The CPU emulation found such a dependence in dft_rows_1e6 - a shuffle that read lanes which do
    not exist - that the single-batch tests could not see.)"""
    n = prm.get("n", 1 << 20)
    a, b = oracle.lcg_bytes(50, 3, n), oracle.lcg_bytes(60, 2, n)
    if not half_is_5_smooth(n):
        prm = dict(prm, test_mask=sts.ALL_TESTS & ~(1 << 7))     # the Bluestein FFT is judged by the last test of this file
    with sts.Engine(max_streams=3, pipeline_depth=2, **prm) as eng:
        first = eng.run(a, 3, 0)
        eng.run(b, 2, 3)
        again = eng.run(a, 3, 5)
        mask = eng.layout.enabled_mask
    with sts.Engine(max_streams=3, **prm) as fresh:
        alone = fresh.run(a, 3, 0)
    for x, y, z in zip(first, again, alone):
        assert (x == y).all() and (x == z).all()
    assert mask != 0


@pytest.mark.parametrize("test,prm", [
    (14, dict(n=1 << 24, serial_m=21)),                  # Serial at the engine's largest block length (64 histogram slices)
    (11, dict(n=1 << 26, apen_m=20)),                    # ApEn at the engine's largest block length (H = 21), 8 MiB streams
    (10, dict(n=1010 * (1 << 15) * 15)),                 # Universal L = 15, the largest table the engine takes (n = 496 Mbit)
])
def test_engine_limits_against_oracle(sts, oracle, test, prm):
    """the largest parameters the engine accepts: one stream each, against the oracle"""
    base = dict(n=1 << 20, block_frequency_M=16384, non_overlapping_m=9, overlapping_m=9, apen_m=10, serial_m=16,
                linear_complexity_M=500)
    base.update(prm)
    mask = 1 << test
    raw = oracle.lcg_bytes(3, 1, base["n"])
    with sts.Engine(max_streams=1, test_mask=mask, **base) as eng:
        assert eng.layout.enabled_mask == mask == oracle.layout(oracle.make_params(test_mask=mask, **base)).enabled_mask
        pv, st, w = eng.run(raw, 1)
        compare_with_oracle(sts, oracle, eng, mask, base, raw, 1, pv, st, w)


@pytest.mark.parametrize("n", [
    1000008,            # n/2 = 2^2 3^2 17 19 43, M = 2^20: also the golden case of the reference binary (lcg_n1000008_4)
    7 * (1 << 17),      # M = 2^20 from 2N - 1 just below it
    8072,               # odd log2 M (one radix-2 pass), N = 4 x 1009
    8 * 1031,           # N = 4 x 1031, M = 2^14 (radix 4 only)
])
def test_dft_bluestein_path(sts, oracle, n):
    """the DFT test on stream lengths whose half is not 5-smooth (k_dft_any.cu): N_1 and the p-value against the oracle,
    and at n = 1000008 against the reference binary's golden values"""
    assert sts.dft_supported(n) and not half_is_5_smooth(n)
    streams, mask = 4, 1 << 7
    prm = dict(n=n, block_frequency_M=16384, non_overlapping_m=9, overlapping_m=9, apen_m=10, serial_m=16,
               linear_complexity_M=500)
    g = load_golden("lcg_n1000008_4") if n == 1000008 else None
    raw = golden_input(g, oracle) if g is not None else oracle.lcg_bytes(3, streams, n)
    with sts.Engine(max_streams=streams, test_mask=mask, **prm) as eng:
        assert eng.layout.enabled_mask == mask
        pv, st, w = eng.run(raw, streams)
        compare_with_oracle(sts, oracle, eng, mask, prm, raw, streams, pv, st, w)
        if g is not None:
            assert (st[:, eng.layout.st_off[7]] == g["dft_N1"]).all()
            assert rel_err(pv[:, eng.layout.pv_off[7]], g["pv_7"].reshape(-1)).max() <= PV_RTOL
