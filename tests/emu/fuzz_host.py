#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: the C host driver linked against the emulated engine (tests/emu_util.py: stsb200_emu) against the
UNMODIFIED reference binary (oracle/_ref/sts) end to end on random command lines: -S n, -i, -P parameters, -t subsets,
-B batch sizes; the bodies of the two result.txt files must be identical, and so must the p-values of the .pvalues files
within 1e-9.  Needs oracle/_ref (build container only).
usage: python tests/emu/fuzz_host.py first_seed last_seed"""
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_util import emu_host_driver  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "sts")
if os.environ.get("STS_FUZZ_GPU", "") not in ("", "0"):	# the product's own host driver on a B200
    EXE = os.path.join(ROOT, "sts_b200", "host", "stsb200")
else:
    EXE = emu_host_driver()


def body(path):
    txt = open(path).read()
    return txt[txt.index("- - - -"):]


def pvalues(path):
    buf = open(path, "rb").read()
    off, out = 0, {}
    while off < len(buf):
        t, cnt = struct.unpack_from("<qq", buf, off)
        off += 16
        out[t] = np.frombuffer(buf, dtype="<f8", count=cnt, offset=off).copy()
        off += 8 * cnt
    return out


bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1000, 4104, 38912, 65536, 131072, 400000, 1000000, 1 << 20]))
    it = int(rng.randint(2, 7))
    P = {}
    if rng.randint(0, 2): P[1] = int(rng.choice([20, 128, 1000, 16384, n // 3 + 1]))
    if rng.randint(0, 2): P[2] = int(rng.choice([8, 9, 10]))
    if rng.randint(0, 2): P[4] = int(rng.randint(1, 12))
    if rng.randint(0, 2): P[5] = int(rng.randint(3, 16))
    if rng.randint(0, 2): P[6] = int(rng.choice([500, 777, 1023]))
    tests = sorted(set(int(t) for t in rng.randint(1, 16, size=int(rng.randint(3, 16)))))
    if n < 1000000:
        tests = [t for t in tests if t != 7 or n in (65536, 131072)] or [1]     # DFT lengths the oracle-free reference FFT is fast on
    tmp = tempfile.mkdtemp(prefix="fuzz_host_")
    try:
        data = os.path.join(tmp, "d.bin")
        O.lcg_bytes(int(rng.randint(0, 50)), it, n).tofile(data)
        common = ["-S", str(n), "-i", str(it), "-F", "r", "-t", ",".join(map(str, tests))]
        if P:
            common += ["-P", ",".join("%d=%d" % kv for kv in sorted(P.items()))]
        ok, why = True, ""
        for mode in ("b", "i"):
            a = subprocess.run([EXE, "-m", mode, "-B", str(int(rng.randint(1, 5))), "-w", os.path.join(tmp, "e" + mode)] + common + [data],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            b = subprocess.run([REF, "-v", "0", "-T", "1", "-m", mode, "-w", os.path.join(tmp, "r" + mode)] + common + [data],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if a.returncode != b.returncode:
                # both refuse (e.g. every selected test disabled itself) or both run
                ok, why = (a.returncode != 0) == (b.returncode != 0), "exit %d vs %d: %s" % (a.returncode, b.returncode, a.stdout[-300:])
                break
            if a.returncode != 0:
                break
            if mode == "b":
                if body(os.path.join(tmp, "eb", "result.txt")) != body(os.path.join(tmp, "rb", "result.txt")):
                    ok, why = False, "result.txt differs"
            else:
                name = "sts.0000.%d.%d.pvalues" % (it, n)
                pa, pb = pvalues(os.path.join(tmp, "ei", name)), pvalues(os.path.join(tmp, "ri", name))
                if sorted(pa) != sorted(pb):
                    ok, why = False, "tests in .pvalues %s vs %s" % (sorted(pa), sorted(pb))
                else:
                    for t in pa:
                        with np.errstate(all="ignore"):
                            rel = np.abs(pa[t] - pb[t]) / np.abs(pb[t])
                        rel[pa[t] == pb[t]] = 0
                        if pa[t].shape != pb[t].shape or not (rel <= 1e-9).all():
                            ok, why = False, "p-values of test %d differ" % t
        print("seed", seed, "n", n, "it", it, "tests", tests, "P", P, "OK" if ok else "MISMATCH: " + why, flush=True)
        bad += not ok
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
print("bad", bad)
sys.exit(1 if bad else 0)
