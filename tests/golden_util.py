"""helpers shared by the oracle-vs-golden (CPU) and engine-vs-oracle/golden (GPU) tests"""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")


def golden_cases():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    # assess_*: assessment reports (make_assess_golden.py); ascii_*: -F a readings (make_ascii_golden.py)
    return [x for x in names if not x.startswith(("assess_", "ascii_"))]


def ascii_lines(bits, line):
    """the bits as '0' / '1' characters (uint8) with a line break after every `line` digits, and one at the end"""
    ch = (np.asarray(bits, dtype=np.uint8) + ord("0")).astype(np.uint8)
    rows = -(-ch.size // line)
    out = np.full((rows, line + 1), ord("\n"), dtype=np.uint8)
    flat = np.full(rows * line, ord(" "), dtype=np.uint8)
    flat[:ch.size] = ch
    out[:, :line] = flat.reshape(rows, line)
    return out.reshape(-1)


def ascii_seek_rule_streams(text, n, streams):
    """parseBitsASCIIInput (utilities.c:1682-1697): iteration k = the first n digits at or after character k n,
    packed MSB first; returns (packed [streams][n / 8], number of complete streams)"""
    text = np.asarray(text, dtype=np.uint8)
    out = np.zeros((streams, n // 8), dtype=np.uint8)
    complete = 0
    for k in range(streams):
        w = text[k * n:]
        d = w[(w == ord("0")) | (w == ord("1"))][:n] - ord("0")
        if d.size < n:
            break
        out[k] = np.packbits(d)
        complete += 1
    return out, complete


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def golden_params(g):
    return dict(n=int(g["n"]), block_frequency_M=int(g["param_block_frequency_M"]),
                non_overlapping_m=int(g["param_non_overlapping_m"]), overlapping_m=int(g["param_overlapping_m"]),
                apen_m=int(g["param_apen_m"]), serial_m=int(g["param_serial_m"]),
                linear_complexity_M=int(g["param_linear_complexity_M"]))


def pattern_streams(n, streams):
    """the degenerate inputs of tests/golden/make_golden.py (kept in sync by test_oracle_golden)"""
    nb = n // 8
    out = np.zeros((streams, nb), dtype=np.uint8)
    rng = np.random.RandomState(12345)
    for s in range(streams):
        k = s % 6
        if k == 1:
            out[s, :] = 0xFF
        elif k == 2:
            out[s, :] = 0x55
        elif k == 3:
            out[s, :] = 0x33
        elif k == 4:
            out[s, nb // 2] = 0x10
        elif k == 5:
            b = rng.randint(0, 256, size=(nb, 3)).astype(np.uint8)
            out[s, :] = b[:, 0] | b[:, 1] | b[:, 2]
    return out.reshape(-1)


def golden_input(g, O):
    n, streams, first = int(g["n"]), int(g["streams"]), int(g["first"])
    if str(g["kind"]) == "lcg":
        return O.lcg_bytes(first, streams, n)
    return pattern_streams(n, streams)


def rel_err(a, b):
    """elementwise relative error with exact matches (incl. 0 == 0, NON_P_VALUE) counted as 0"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.abs(b)
    r[a == b] = 0.0
    r[np.isnan(r)] = np.inf
    return r


def check_against_golden(g, pv, st, w, lay, pv_rtol, dft_exact=True):
    """compare records (pvals, istats, wcount with layout `lay`) with one golden case; returns list of failures"""
    bad = []
    streams = int(g["streams"])
    if int(g["enabled_mask"]) != lay.enabled_mask:
        bad.append("enabled mask %#x != golden %#x" % (lay.enabled_mask, int(g["enabled_mask"])))
        return bad

    def ST(t, lo, hi=None):
        o = lay.st_off[t]
        return st[:streams, o + lo:o + (hi if hi is not None else lo + 1)]

    for t in range(1, 16):
        if not (lay.enabled_mask >> t) & 1:
            continue
        ref = g["pv_%d" % t]
        mine = pv[:streams, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
        if mine.shape != ref.shape:
            bad.append("test %d: p-value shape %s != %s" % (t, mine.shape, ref.shape))
            continue
        r = rel_err(mine, ref)
        if r.max() > pv_rtol:
            i = np.unravel_index(np.argmax(r), r.shape)
            bad.append("test %d: p-value rel err %.3g at %s (%r vs %r)" % (t, r.max(), i, mine[i], ref[i]))
    en = lay.enabled_mask

    def eq(name, a, b):
        a = np.asarray(a)
        b = np.asarray(b)
        if a.shape != b.shape or not (a == b).all():
            bad.append("%s mismatch: %s vs %s" % (name, a.ravel()[:12].tolist(), b.ravel()[:12].tolist()))

    if en >> 1 & 1:
        eq("freq_S_n", ST(1, 0)[:, 0], g["freq_S_n"])
    if en >> 3 & 1:
        eq("cusum_z", ST(3, 3, 5), g["cusum_z"])
    if en >> 4 & 1:
        v = g["runs_V_n"]
        m = v >= 0
        eq("runs_V_n", ST(4, 1)[:, 0][m], v[m])
        eq("runs_possible", ST(4, 2)[:, 0], m.astype(np.int64))
    if en >> 5 & 1:
        eq("longest_run_counts", ST(5, 0, 7), g["longest_run_counts"])
    if en >> 6 & 1:
        eq("rank_F", ST(6, 0, 3), g["rank_F"])
    if en >> 7 & 1 and dft_exact:
        eq("dft_N1", ST(7, 0)[:, 0], g["dft_N1"])
    if en >> 8 & 1:
        eq("nonov_W", w[:streams].reshape(streams, lay.ntemplates, 8), g["nonov_W"])
    if en >> 9 & 1:
        eq("overlap_v", ST(9, 0, 6), g["overlap_v"])
    if en >> 12 & 1:
        eq("rex_J", ST(12, 0)[:, 0], g["rex_J"])
    if en >> 13 & 1:
        eq("rexv_J", ST(13, 0)[:, 0], g["rexv_J"])
        xi = g["rexv_xi"]
        m = xi[:, 0] >= 0
        eq("rexv_xi", ST(13, 1, 19)[m], xi[m])
    if en >> 15 & 1:
        eq("lc_v", ST(15, 0, 7), g["lc_v"])
    return bad
