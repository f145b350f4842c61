"""The real drop-in (integration/): the reference's own host program - sts.c, parse_args.c, driver.c with its
testDriver[] table and iterate(), every X_init / X_print / X_metrics, write_p_val_to_file, read_from_p_val_file,
metrics() compiled from /root/reference in place - with the fifteen X_iterate bodies and invokeTestSuite renamed away
and defined by integration/glue.c over the engine.  Compared with the unmodified reference binary (oracle/_ref/sts)
on the same files: result.txt, the .pvalues file, every <Test>/stats.txt / results.txt / data*.txt of -s, the
-m i / -m a flow, -j, -F a, stdin, -O.

The bodies are written once and run twice: `-m gpu` on a B200 against integration/_ref/sts_gpu (the product), and in
the CPU suite (tests/test_emu_integration.py) against the same glue linked with the emulated engine of tests/emu."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from test_gpu_host_driver import read_pvalues, result_body, run, table_lines, lcg8  # noqa: F401

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STS_GPU = os.path.join(ROOT, "integration", "_ref", "sts_gpu")
REF = os.path.join(ROOT, "oracle", "_ref", "sts")
TESTS = ["Frequency", "BlockFrequency", "CumulativeSums", "Runs", "LongestRun", "Rank", "DFT", "NonOverlappingTemplate",
         "OverlappingTemplate", "Universal", "ApproximateEntropy", "RandomExcursions", "RandomExcursionsVariant", "Serial",
         "LinearComplexity"]
PV_RTOL = 1e-9


def need_binaries():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    assert os.path.exists(STS_GPU), "integration/_ref/sts_gpu not built (make -C integration)"


def same_pvalues(a, b, cusum_atol=1e-13):
    pa, pb = read_pvalues(a), read_pvalues(b)
    assert sorted(pa) == sorted(pb)
    for t in pa:
        assert pa[t].shape == pb[t].shape, "test %d" % t
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.abs(pa[t] - pb[t]) / np.abs(pb[t])
        r[pa[t] == pb[t]] = 0.0
        if t == 3:          # 1 - sum + sum (cusum.c:369): absolute rounding noise of O(1) sums
            r[np.abs(pa[t] - pb[t]) <= cusum_atol] = 0.0
        assert r.max() <= PV_RTOL, "test %d: %g" % (t, r.max())


def test_no_cpu_test_code_is_linked():
    """the renamed CPU iterate bodies, the pthread dispatch and the CPU FFT entry are not in the binary at all"""
    assert os.path.exists(STS_GPU), "integration/_ref/sts_gpu not built (make -C integration)"
    syms = subprocess.run(["nm", STS_GPU], stdout=subprocess.PIPE, text=True, check=True).stdout
    for s in ("_iterate_cpu", "invokeTestSuite_cpu", "testBits", "parseBitsBinaryInput", "copyBitsToEpsilon", "fftw_execute"):
        assert s not in syms, s
    for s in ("Frequency_iterate", "LinearComplexity_iterate", "invokeTestSuite", "testDriver", "Frequency_print", "metrics",
              "write_p_val_to_file", "read_from_p_val_file"):
        assert s in syms, s


def test_result_txt_and_pvalues_against_reference(tmp_path, oracle):
    """sts_gpu -i 32 -F r and sts -i 32 -F r -T 1 (BASELINE config 1): identical result.txt; -m i: the same .pvalues"""
    need_binaries()
    p = str(tmp_path / "lcg32.bin")
    oracle.lcg_bytes(0, 32).tofile(p)
    env_batch = dict(os.environ, STS_GPU_BATCH="12")                 # ragged batches 12 + 12 + 8
    g, r = str(tmp_path / "gpu"), str(tmp_path / "ref")
    subprocess.run([STS_GPU, "-v", "0", "-i", "32", "-F", "r", "-w", g, p], check=True, env=env_batch, stdout=subprocess.DEVNULL)
    run([REF, "-v", "0", "-i", "32", "-F", "r", "-T", "1", "-w", r, p])
    assert result_body(os.path.join(g, "result.txt")) == result_body(os.path.join(r, "result.txt"))
    gi, ri = str(tmp_path / "gpu_i"), str(tmp_path / "ref_i")
    run([STS_GPU, "-v", "0", "-m", "i", "-i", "32", "-F", "r", "-w", gi, p])
    run([REF, "-v", "0", "-m", "i", "-i", "32", "-F", "r", "-T", "1", "-w", ri, p])
    same_pvalues(os.path.join(gi, "sts.0000.32.1048576.pvalues"), os.path.join(ri, "sts.0000.32.1048576.pvalues"))


def test_iterate_only_jobs_then_assess_only(lcg8, tmp_path):
    """the reference's distributed flow: -m i -j 0 / -j 1 on halves of the file, then -m a -d over both files;
    every combination of who iterates and who assesses gives the reference's result.txt"""
    need_binaries()
    dirs = {}
    for name, exe, extra in (("gpu", STS_GPU, []), ("ref", REF, ["-T", "1"])):
        d = str(tmp_path / name)
        for j in ("0", "1"):
            run([exe, "-v", "0", "-m", "i", "-i", "4", "-j", j, "-F", "r", "-w", d] + extra + [lcg8])
        dirs[name] = d
    for j in ("0000", "0001"):
        same_pvalues(os.path.join(dirs["gpu"], "sts.%s.4.1048576.pvalues" % j), os.path.join(dirs["ref"], "sts.%s.4.1048576.pvalues" % j))
    bodies = []
    for who, exe in (("gpu", STS_GPU), ("ref", REF)):
        for src in ("gpu", "ref"):
            w = str(tmp_path / ("assess_%s_on_%s" % (who, src)))
            run([exe, "-v", "0", "-m", "a", "-d", dirs[src], "-i", "8", "-w", w])
            bodies.append(result_body(os.path.join(w, "result.txt")))
    assert bodies[0] == bodies[1] == bodies[2] == bodies[3]


def compare_dash_s(tmp_path, file, iters, extra=()):
    g, r = str(tmp_path / "gpu_s"), str(tmp_path / "ref_s")
    run([STS_GPU, "-v", "0", "-s", "-i", str(iters), "-F", "r", "-w", g] + list(extra) + [file])
    run([REF, "-v", "0", "-s", "-T", "1", "-i", str(iters), "-F", "r", "-w", r] + list(extra) + [file])
    assert result_body(os.path.join(g, "result.txt")) == result_body(os.path.join(r, "result.txt"))
    seen = 0
    for t in TESTS:
        rd = os.path.join(r, t)
        if not os.path.isdir(rd):
            continue
        for f in sorted(os.listdir(rd)):
            a, b = os.path.join(g, t, f), os.path.join(rd, f)
            assert os.path.exists(a), "%s/%s missing" % (t, f)
            if not filecmp.cmp(a, b, shallow=False):
                la, lb = open(a).read().splitlines(), open(b).read().splitlines()
                bad = [(i, x, y) for i, (x, y) in enumerate(zip(la, lb)) if x != y][:5]
                raise AssertionError("%s/%s differs (%d vs %d lines): %s" % (t, f, len(la), len(lb), bad))
            seen += 1
    return seen


def test_dash_s_stats_results_and_data_files_identical(lcg8, tmp_path):
    """-s: every <Test>/stats.txt (the per-iteration intermediate values of each private stats struct, the last-cycle
    visit counts of RandomExcursions included), results.txt and data*.txt is byte-identical to the reference's"""
    need_binaries()
    assert compare_dash_s(tmp_path, lcg8, 8) >= 190         # 15 stats.txt + 15 results.txt + the data*.txt of the five multi-value tests


def test_dash_s_on_degenerate_streams(tmp_path):
    """six degenerate streams (all 0, all 1, 0101.., 0011.., one-hot, 7/8 biased): tests that are not possible
    (Runs, the excursions) and p-values of 0 go through the same -s files"""
    need_binaries()
    n = 1 << 20
    rows = [np.zeros(n // 8, np.uint8), np.full(n // 8, 0xFF, np.uint8), np.full(n // 8, 0x55, np.uint8),
            np.full(n // 8, 0x33, np.uint8), np.zeros(n // 8, np.uint8), np.full(n // 8, 0xFE, np.uint8)]
    rows[4] = rows[4].copy()
    rows[4][n // 16] = 0x10
    p = str(tmp_path / "patterns.bin")
    np.concatenate(rows).tofile(p)
    compare_dash_s(tmp_path, p, 6)


def test_non_default_parameters_and_legacy_output(lcg8, tmp_path):
    """-P and -t on the unmodified parse_args, -O: freq.txt (BITSREAD lines) and finalAnalysisReport.txt"""
    need_binaries()
    args = ["-v", "0", "-O", "-i", "8", "-F", "r", "-t", "1,2,4,8,11,14,15", "-P", "1=20000,2=10,4=8,5=12,6=1000"]
    g, r = str(tmp_path / "gpu"), str(tmp_path / "ref")
    run([STS_GPU] + args + ["-w", g, lcg8])
    run([REF] + args + ["-T", "1", "-w", r, lcg8])
    assert table_lines(os.path.join(g, "finalAnalysisReport.txt")) == table_lines(os.path.join(r, "finalAnalysisReport.txt"))
    fa = [l for l in open(os.path.join(g, "freq.txt")) if "BITSREAD" in l]
    fb = [l for l in open(os.path.join(r, "freq.txt")) if "BITSREAD" in l]
    assert len(fa) == 8 and fa == fb


def test_ascii_and_stdin_input(tmp_path, oracle):
    """-F a from a file with line breaks (text packed on the device, the reference's per-iteration seek), raw bytes on stdin"""
    need_binaries()
    raw = oracle.lcg_bytes(0, 2)
    bits = np.unpackbits(raw)
    txt = "\n".join("".join("01"[b] for b in bits[i:i + 80]) for i in range(0, bits.size, 80)) + "\n"
    ap = str(tmp_path / "ascii.txt")
    open(ap, "w").write(txt)
    g, r = str(tmp_path / "ga"), str(tmp_path / "ra")
    run([STS_GPU, "-v", "0", "-m", "i", "-i", "2", "-F", "a", "-t", "1,3,4,9,13", "-w", g, ap])
    run([REF, "-v", "0", "-m", "i", "-i", "2", "-F", "a", "-T", "1", "-t", "1,3,4,9,13", "-w", r, ap])
    same_pvalues(os.path.join(g, "sts.0000.2.1048576.pvalues"), os.path.join(r, "sts.0000.2.1048576.pvalues"))
    gs, rs = str(tmp_path / "gs"), str(tmp_path / "rs")
    for exe, w, extra in ((STS_GPU, gs, []), (REF, rs, ["-T", "1"])):
        q = subprocess.run([exe, "-v", "0", "-m", "i", "-i", "2", "-F", "r", "-t", "1,6,10", "-w", w] + extra + ["-"],
                           input=raw.tobytes(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert q.returncode == 0, q.stdout[-1500:]
    same_pvalues(os.path.join(gs, "sts.0000.2.1048576.pvalues"), os.path.join(rs, "sts.0000.2.1048576.pvalues"))


def test_short_raw_file_is_fatal_like_the_reference(tmp_path, oracle):
    need_binaries()
    p = str(tmp_path / "short.bin")
    oracle.lcg_bytes(0, 2)[:-100].tofile(p)
    codes = []
    for exe, extra in ((STS_GPU, []), (REF, ["-T", "1"])):
        q = subprocess.run([exe, "-v", "0", "-i", "2", "-F", "r", "-w", str(tmp_path / os.path.basename(exe))] + extra + [p],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert "EOF" in q.stdout
        codes.append(q.returncode)
    assert codes[0] == codes[1] == 226
