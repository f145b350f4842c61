"""GPU (-m gpu): the C host driver (sts_b200/host/stsb200, the reference's main() sequence for the
hot path) end to end: same flags, same .pvalues bytes layout, and - when the prebuilt reference binary
travelled with the repo (oracle/_ref/sts) - the SAME result.txt as the reference run on the same file
(north star: "matching result.txt")."""
import os
import struct
import subprocess

import numpy as np
import pytest

from golden_util import load_golden, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STSB200 = os.path.join(ROOT, "sts_b200", "host", "stsb200")
REF = os.path.join(ROOT, "oracle", "_ref", "sts")


def read_pvalues(path):
    buf = open(path, "rb").read()
    off, out = 0, {}
    while off < len(buf):
        t, cnt = struct.unpack_from("<qq", buf, off)
        off += 16
        out[t] = np.frombuffer(buf, dtype="<f8", count=cnt, offset=off).copy()
        off += 8 * cnt
    return out


def result_body(path):
    txt = open(path).read()
    return txt[txt.index("- - - -"):]        # the header names the data source, the rest must be identical


@pytest.fixture(scope="module")
def lcg8(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("lcg")
    p = str(d / "lcg8.bin")
    oracle.lcg_bytes(0, 8).tofile(p)
    return p


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, "%s\n%s" % (" ".join(cmd), r.stdout[-2000:])
    return r.stdout


def test_pvalues_file_matches_reference_golden(lcg8, tmp_path):
    assert os.path.exists(STSB200), "host driver not built (make -C sts_b200)"
    w = str(tmp_path / "w")
    run([STSB200, "-m", "i", "-i", "8", "-F", "r", "-B", "3", "-w", w, lcg8])      # -B 3: ragged batches 3+3+2
    pv = read_pvalues(os.path.join(w, "sts.0000.8.1048576.pvalues"))
    g = load_golden("lcg_default_8")
    assert sorted(pv) == list(range(1, 16))
    for t in range(1, 16):
        ref = g["pv_%d" % t]
        mine = pv[t].reshape(ref.shape)          # iteration-major, like the reference with -T 1
        assert rel_err(mine, ref).max() <= 1e-9, "test %d" % t


def test_result_txt_identical_to_reference(lcg8, tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    w = str(tmp_path / "gpu")
    run([STSB200, "-m", "i", "-i", "8", "-F", "r", "-w", w, lcg8])
    run([REF, "-v", "0", "-m", "a", "-d", w, "-i", "8", "-w", str(tmp_path / "assess")])
    run([REF, "-v", "0", "-i", "8", "-T", "8", "-F", "r", "-w", str(tmp_path / "ref"), lcg8])
    a = result_body(str(tmp_path / "assess" / "result.txt"))
    b = result_body(str(tmp_path / "ref" / "result.txt"))
    assert a == b


def table_lines(path):
    import re
    return [l.rstrip() for l in open(path) if re.search(r"\d+/\d+\s", l) and "generator" not in l]


def test_legacy_assessment_table_identical_to_reference(tmp_path, oracle):
    """stsb200 -O -m b: the uniformity / proportion table (tallied on the GPU, stsgpu_assess_*) is, line for
    line, the table of the reference's own finalAnalysisReport.txt for the same 32 streams"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    p = str(tmp_path / "lcg32.bin")
    oracle.lcg_bytes(0, 32).tofile(p)
    run([STSB200, "-O", "-m", "b", "-i", "32", "-F", "r", "-B", "12", "-w", str(tmp_path / "gpu"), p])
    run([REF, "-v", "0", "-O", "-i", "32", "-F", "r", "-w", str(tmp_path / "ref"), p])
    a = table_lines(str(tmp_path / "gpu" / "finalAnalysisReport.txt"))
    b = table_lines(str(tmp_path / "ref" / "finalAnalysisReport.txt"))
    assert len(a) == len(b) == 188
    assert a == b


def test_jobnum_and_ascii_input(lcg8, tmp_path, oracle):
    w = str(tmp_path / "j")
    run([STSB200, "-m", "i", "-i", "4", "-j", "1", "-t", "1,3,8,15", "-w", w, lcg8])   # second half of the file
    pv = read_pvalues(os.path.join(w, "sts.0001.4.1048576.pvalues"))
    g = load_golden("lcg_default_8")
    assert sorted(pv) == [1, 3, 8, 15]
    for t in (1, 3, 8, 15):
        ref = g["pv_%d" % t][4:]
        assert rel_err(pv[t].reshape(ref.shape), ref).max() <= 1e-9
    # -F a: the same first stream as ASCII 0/1 characters with line breaks
    bits = np.unpackbits(oracle.lcg_bytes(0, 1))
    txt = "\n".join("".join("01"[b] for b in bits[i:i + 64]) for i in range(0, bits.size, 64))
    ap = str(tmp_path / "ascii.txt")
    open(ap, "w").write(txt + "\n")
    wa = str(tmp_path / "a")
    run([STSB200, "-m", "i", "-i", "1", "-F", "a", "-w", wa, ap])
    pa = read_pvalues(os.path.join(wa, "sts.0000.1.1048576.pvalues"))
    for t in range(1, 16):
        ref = g["pv_%d" % t][:1]
        assert rel_err(pa[t].reshape(ref.shape), ref).max() <= 1e-9


def test_distributed_mode_script(lcg8, tmp_path):
    """tools/run_distributed.sh: the reference's `-m i -j K` / `-m a -d DIR` flow with one job per process; here two
    jobs on the one GPU that is guaranteed to exist (jobs only differ in -j and -D), assessed by the reference"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    d = str(tmp_path / "dist")
    os.makedirs(d)
    for k in range(2):
        run([STSB200, "-m", "i", "-D", "0", "-j", str(k), "-i", "4", "-F", "r", "-w", d, lcg8])
    run([REF, "-v", "0", "-m", "a", "-d", d, "-i", "8", "-w", str(tmp_path / "assess")])
    run([REF, "-v", "0", "-i", "8", "-T", "8", "-F", "r", "-w", str(tmp_path / "ref"), lcg8])
    assert result_body(str(tmp_path / "assess" / "result.txt")) == result_body(str(tmp_path / "ref" / "result.txt"))


def test_short_file_is_fatal(tmp_path):
    p = str(tmp_path / "short.bin")
    open(p, "wb").write(b"\x00" * 1000)
    r = subprocess.run([STSB200, "-i", "1", "-w", str(tmp_path / "s"), p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 213                   # utilities.c:1789-1791: a short raw read is fatal


def test_ascii_file_with_line_breaks_matches_reference(tmp_path, oracle):
    """-F a through the GPU packer: in a file with line breaks the reference's per-iteration seek makes consecutive
    streams overlap; stsb200 must produce the reference's p-values, not those of a sequential reader.  Also -j 1."""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    bits = np.unpackbits(oracle.lcg_bytes(0, 7))
    txt = "\n".join("".join("01"[b] for b in bits[i:i + 80]) for i in range(0, bits.size, 80)) + "\n"
    ap = str(tmp_path / "ascii.txt")
    open(ap, "w").write(txt)
    for job in ("0", "1"):
        wg, wr = str(tmp_path / ("g" + job)), str(tmp_path / ("r" + job))
        run([STSB200, "-m", "i", "-i", "3", "-j", job, "-B", "2", "-F", "a", "-w", wg, ap])
        run([REF, "-v", "0", "-m", "i", "-T", "1", "-i", "3", "-j", job, "-F", "a", "-w", wr, ap])
        name = "sts.000%s.3.1048576.pvalues" % job
        a, b = read_pvalues(os.path.join(wg, name)), read_pvalues(os.path.join(wr, name))
        assert sorted(a) == sorted(b) == list(range(1, 16))
        for t in range(1, 16):
            assert a[t].shape == b[t].shape
            assert rel_err(a[t], b[t]).max() <= 1e-9, "test %d" % t


def test_stdin_input(lcg8, tmp_path):
    """datafile "-": standard input is read in sequence, never seeked (utilities.c:1500-1502)"""
    w = str(tmp_path / "stdin")
    with open(lcg8, "rb") as f:
        r = subprocess.run([STSB200, "-m", "i", "-i", "2", "-t", "1,2,4", "-F", "r", "-w", w, "-"], stdin=f,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout[-2000:]
    pv = read_pvalues(os.path.join(w, "sts.0000.2.1048576.pvalues"))
    g = load_golden("lcg_default_8")
    for t in (1, 2, 4):
        ref = g["pv_%d" % t][:2]
        assert rel_err(pv[t].reshape(ref.shape), ref).max() <= 1e-9


def test_result_txt_of_mode_b_identical_to_reference(tmp_path, oracle):
    """stsb200 -m b writes result.txt itself (verdicts from the GPU tallies): same body as the reference's"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    p = str(tmp_path / "lcg32.bin")
    oracle.lcg_bytes(0, 32).tofile(p)
    run([STSB200, "-m", "b", "-i", "32", "-F", "r", "-B", "20", "-w", str(tmp_path / "gpu"), p])
    run([REF, "-v", "0", "-i", "32", "-F", "r", "-w", str(tmp_path / "ref"), p])
    a, b = result_body(str(tmp_path / "gpu" / "result.txt")), result_body(str(tmp_path / "ref" / "result.txt"))
    assert "tests passed successfully both the analyses" in a
    assert a == b
    head = open(str(tmp_path / "gpu" / "result.txt")).read().split("- - - -")[0]
    assert "A total of 188 tests" in head and "32 bitstreams of 1048576 bits" in head and p in head


def test_assess_only_mode_streams_pvalues_files(lcg8, tmp_path):
    """stsb200 -m a -d DIR: the .pvalues files of two -m i jobs, streamed through the GPU tallies, give the
    result.txt of the reference's own sts -m a on the same directory (with and without -i: the bins quirk)"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    d = str(tmp_path / "pv")
    os.makedirs(d)
    for k in range(2):
        run([STSB200, "-m", "i", "-j", str(k), "-i", "4", "-F", "r", "-w", d, lcg8])
    open(os.path.join(d, "sts.0000.4.4096.pvalues"), "wb").write(b"")        # other n: must be ignored
    for extra in (["-i", "8"], []):
        tag = "i" if extra else "noi"
        run([STSB200, "-m", "a", "-d", d, "-w", str(tmp_path / ("gpu_" + tag))] + extra)
        run([REF, "-v", "0", "-m", "a", "-d", d, "-w", str(tmp_path / ("ref_" + tag))] + extra)
        a = open(str(tmp_path / ("gpu_" + tag) / "result.txt")).read()
        b = open(str(tmp_path / ("ref_" + tag) / "result.txt")).read()
        assert "8 bitstreams of 1048576 bits" in a
        assert a == b                                # whole file: same source line (--SOME_FILE--) and directory


def test_dash_s_results_and_data_files_identical_to_reference(lcg8, tmp_path):
    """-s: <TestName>/results.txt and data*.txt (the p-value files of every X_print, including the reference's file
    naming: data.txt for 2 and 8 partitions, data1..18, data01..148) byte for byte; stats.txt is not produced"""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/sts not shipped")
    wg, wr = str(tmp_path / "g"), str(tmp_path / "r")
    run([STSB200, "-s", "-m", "b", "-i", "3", "-F", "r", "-w", wg, lcg8])
    run([REF, "-v", "0", "-s", "-T", "1", "-i", "3", "-F", "r", "-w", wr, lcg8])
    seen = 0
    for root, _, files in os.walk(wr):
        for f in files:
            if f == "results.txt" or f.startswith("data"):
                rel = os.path.relpath(os.path.join(root, f), wr)
                mine = os.path.join(wg, rel)
                assert os.path.exists(mine), rel
                assert open(mine).read() == open(os.path.join(root, f)).read(), rel
                seen += 1
    assert seen == 15 + 1 + 148 + 1 + 18 + 1       # results.txt x 15, cusum, templates, excursions, variant, serial


def test_generator_1_on_the_device(tmp_path, oracle):
    """-g 1: the linear congruential generator of tools/generators made on the GPU (skip-ahead) is the file the
    generator writes - same p-values as the golden file run, -j jobs start at their own streams, and -m b keeps no
    p-values yet writes the same result.txt as a run on the file"""
    g = load_golden("lcg_default_8")
    w = str(tmp_path / "gen")
    run([STSB200, "-g", "1", "-m", "i", "-i", "8", "-B", "3", "-w", w])
    pv = read_pvalues(os.path.join(w, "sts.0000.8.1048576.pvalues"))
    for t in range(1, 16):
        ref = g["pv_%d" % t]
        assert rel_err(pv[t].reshape(ref.shape), ref).max() <= 1e-9, "test %d" % t
    wj = str(tmp_path / "genj")
    run([STSB200, "-g", "1", "-m", "i", "-j", "1", "-i", "4", "-t", "1,7,15", "-w", wj])
    pj = read_pvalues(os.path.join(wj, "sts.0001.4.1048576.pvalues"))
    for t in (1, 7, 15):
        ref = g["pv_%d" % t][4:]
        assert rel_err(pj[t].reshape(ref.shape), ref).max() <= 1e-9
    p = str(tmp_path / "lcg32.bin")
    oracle.lcg_bytes(0, 32).tofile(p)
    run([STSB200, "-m", "b", "-i", "32", "-F", "r", "-w", str(tmp_path / "file"), p])
    run([STSB200, "-g", "1", "-m", "b", "-i", "32", "-B", "7", "-w", str(tmp_path / "made")])
    assert result_body(str(tmp_path / "file" / "result.txt")) == result_body(str(tmp_path / "made" / "result.txt"))
    assert not [f for f in os.listdir(str(tmp_path / "made")) if f.endswith(".pvalues")]     # sts.c:358: only -m i writes them


@pytest.mark.parametrize("gen", [2, 3, 4, 5, 6, 7, 8, 9])
def test_generators_2_to_9_are_the_files_the_tool_writes(tmp_path, gen):
    """-g 2..9: testing a generator equals testing the file tools/generators writes for it (restated by the oracle, which
    is pinned to the tool's output): 2, 4, 5 made on the GPU, 3, 6, 7, 8, 9 chained on a host core; job 1 of two starts
    at its own streams either way.  Generator 8 is the tool's with its uninitialised accumulator defined as zero."""
    from oracle import generators as OG
    n, its = (1 << 20, 4) if gen not in (6, 7) else (1 << 14, 4)        # 6 and 7 take seconds per megabit stream
    tests = "1,2,3,4" if n == 1 << 20 else "1,3,4"
    p = str(tmp_path / "gen.bin")
    open(p, "wb").write(b"".join(OG.generate(gen, n, its)))
    size = ["-S", str(n)]
    run([STSB200, "-m", "i", "-i", str(its), "-t", tests, "-F", "r", "-w", str(tmp_path / "file")] + size + [p])
    run([STSB200, "-g", str(gen), "-m", "i", "-i", str(its), "-B", "3", "-t", tests, "-w", str(tmp_path / "made")] + size)
    name = "sts.0000.%d.%d.pvalues" % (its, n)
    a, b = read_pvalues(str(tmp_path / "file" / name)), read_pvalues(str(tmp_path / "made" / name))
    assert sorted(a) == sorted(b)
    for t in a:
        assert (a[t] == b[t]).all(), "test %d" % t
    run([STSB200, "-g", str(gen), "-m", "i", "-j", "1", "-i", str(its // 2), "-t", tests, "-w", str(tmp_path / "job1")] + size)
    c = read_pvalues(str(tmp_path / "job1" / ("sts.0001.%d.%d.pvalues" % (its // 2, n))))
    for t in a:
        per = len(a[t]) // its
        assert (c[t] == a[t][per * (its // 2):]).all(), "test %d" % t


def test_ascii_file_against_committed_reference_reading(tmp_path, oracle):
    """the same as test_ascii_file_with_line_breaks_matches_reference, but against the committed fixture
    (tests/golden/ascii_lines80_3.npz, made by the reference binary) so that it also runs where oracle/_ref is absent"""
    from golden_util import GOLDEN_DIR, ascii_lines
    g = np.load(os.path.join(GOLDEN_DIR, "ascii_lines80_3.npz"))
    it = int(g["iterations"])
    ap = str(tmp_path / "ascii.txt")
    ascii_lines(np.unpackbits(oracle.lcg_bytes(0, int(g["source_streams"]))), int(g["line"])).tofile(ap)
    w = str(tmp_path / "w")
    run([STSB200, "-m", "i", "-i", str(it), "-B", "2", "-F", "a", "-w", w, ap])
    pv = read_pvalues(os.path.join(w, "sts.0000.%d.1048576.pvalues" % it))
    for t in range(1, 16):
        ref = g["pv_%d" % t]
        assert rel_err(pv[t].reshape(ref.shape), ref).max() <= 1e-9, "test %d" % t


def test_reports_against_committed_reference_bodies(tmp_path):
    """result.txt of a -m b run and of a -m i x 2 jobs + -m a run, on streams made by -g 1, against the bodies the
    reference binary wrote for the same streams (tests/golden/make_report_golden.py); needs neither oracle/_ref nor
    a data file"""
    gold = os.path.join(ROOT, "tests", "golden")
    run([STSB200, "-g", "1", "-m", "b", "-i", "32", "-B", "13", "-w", str(tmp_path / "b")])
    assert result_body(str(tmp_path / "b" / "result.txt")) == open(os.path.join(gold, "result_lcg32_body.txt")).read()
    d = str(tmp_path / "pv")
    for k in range(2):
        run([STSB200, "-g", "1", "-m", "i", "-j", str(k), "-i", "4", "-w", d])
    run([STSB200, "-m", "a", "-d", d, "-i", "8", "-w", str(tmp_path / "a")])
    assert result_body(str(tmp_path / "a" / "result.txt")) == open(os.path.join(gold, "assess_lcg8_body.txt")).read()
