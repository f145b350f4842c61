#!/usr/bin/env python3
"""Golden fixture for the X_init() self-disable rules (column "init"): which of the 15 tests the
UNMODIFIED reference binary leaves enabled for a stream length and parameter set, and how many p-values per
iteration each one records.

For every case below the reference (oracle/_ref/sts, built by `make -C oracle ref`) runs ONE iteration
(`-T 1 -m i -i 1 -S n [-P ...]`) on the first n bits of the LCG input; the enabled tests and their p-value counts
are read back from its own .pvalues file (utilities.c:1967-2024: int64 test, int64 count, float64[count] per
enabled test).  Result: tests/golden/init_rules.json, a list of
    {"n": ..., "P": {"1": ..., ...}, "enabled_mask": ..., "pv_cnt": [16 ints]}
or {"n": ..., "P": ..., "refused": true} where the reference exits with an error instead of running.

Only runs in the build container (needs oracle/_ref); the tests only read the .json file.
"""
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
STS = os.path.join(ROOT, "oracle", "_ref", "sts")

N20 = 1 << 20
CASES = []


def case(n, **P):
    CASES.append((n, {int(k[1:]): v for k, v in P.items()}))


# stream length thresholds: n >= 100 (frequency.c:102 ...), n >= 128 (longestRunOfOnes.c:261), n >= 1000 (DFT),
# 38 rank matrices, Universal n >= 387840, n >= 10^6 (Overlapping, RandomExcursions*, LinearComplexity)
for n_ in (96, 104, 120, 128, 992, 1000, 2048, 38 * 1024 - 8, 38 * 1024, 65536, 387832, 387840, 400000, 904960,
           999992, 1000000, 1000008, N20, 6 * (1 << 17)):
    case(n_)
# BlockFrequency: M >= 20, M > 0.01 n, N = n / M < 100 (blockFrequency.c:109-129)
for M in (19, 20, 10485, 10486, 10591, 10592, 20000, N20):
    case(N20, P1=M)
case(2048, P1=19)
case(2048, P1=20)
case(2048, P1=21)
# NonOverlapping template length 8..15 at most (nonOverlappingTemplateMatchings.c:211-241); Overlapping only m = 9
for m in (7, 8, 10, 15, 16):
    case(N20, P2=m)
for m in (8, 10):
    case(N20, P3=m)
# ApEn m < floor(log2 n - 5), Serial m < floor(log2 n - 2) (approximateEntropy.c:109, serial.c:112)
for m in (1, 13, 14, 15):
    case(N20, P4=m)
for m in (3, 16, 17, 18):
    case(N20, P5=m)
case(65536, P4=10, P5=13)
case(65536, P4=11, P5=14)
# LinearComplexity 500 <= M <= 5000, N = n / M >= 200 (linearComplexity.c:118-138)
for M in (499, 500, 501, 5000, 5001):
    case(N20, P6=M)
case(1000000, P6=5000)
case(1000000, P6=4999)
# BASELINE config 4 and the same length with defaults (BlockFrequency disables itself)
case(10000000)
case(10000000, P6=5000, P2=10, P4=16, P5=16)


def read_pvalues(path):
    buf = open(path, "rb").read()
    off, out = 0, {}
    while off < len(buf):
        t, cnt = struct.unpack_from("<qq", buf, off)
        off += 16 + 8 * cnt
        out[int(t)] = int(cnt)
    return out


def main():
    from oracle import oracle as O
    nmax = max(n for n, _ in CASES)
    raw = O.lcg_bytes(0, 1, nmax)
    out = []
    tmp = tempfile.mkdtemp(prefix="init_golden_")
    try:
        for n, P in CASES:
            data = os.path.join(tmp, "data.bin")
            raw[:n // 8].tofile(data)
            work = os.path.join(tmp, "work")
            shutil.rmtree(work, ignore_errors=True)
            cmd = [STS, "-T", "1", "-m", "i", "-F", "r", "-i", "1", "-S", str(n), "-w", work]
            if P:
                cmd += ["-P", ",".join("%d=%d" % kv for kv in sorted(P.items()))]
            cmd.append(data)
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            rec = {"n": n, "P": {str(k): v for k, v in sorted(P.items())}}
            pvf = os.path.join(work, "sts.0000.1.%d.pvalues" % n)
            if r.returncode != 0 or not os.path.exists(pvf):
                rec["refused"] = True
                rec["message"] = [l for l in r.stdout.splitlines() if l.strip()][-1][:200] if r.stdout.strip() else ""
            else:
                cnt = read_pvalues(pvf)
                rec["enabled_mask"] = sum(1 << t for t in cnt)
                rec["pv_cnt"] = [cnt.get(t, 0) for t in range(16)]
            out.append(rec)
            print(rec)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "init_rules.json"), "w") as f:
        f.write("[\n" + ",\n".join(json.dumps(r, sort_keys=True) for r in out) + "\n]\n")


if __name__ == "__main__":
    main()
