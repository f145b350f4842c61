#!/usr/bin/env python3
"""Golden vectors for the uniformity / proportion assessment, from the UNMODIFIED reference binary.

Runs oracle/_ref/sts -O (legacy report: 10 uniformity bins, parse_args.c:792-803) in mode b on the first
64 LCG streams and parses finalAnalysisReport.txt, whose table holds, per p-value column in test order,
the ten bin counts, the uniformity p-value (%8.6f) and passCount/sampleCount (frequency.c:700-735).
Stored as tests/golden/assess_lcg_64.npz.  Only runs in the build container (needs oracle/_ref).
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
STS = os.path.join(ROOT, "oracle", "_ref", "sts")
STREAMS = 64


def main():
    from oracle import oracle as O
    tmp = tempfile.mkdtemp(prefix="assess_golden_")
    data = os.path.join(tmp, "data.bin")
    O.lcg_bytes(0, STREAMS).tofile(data)
    subprocess.check_call([STS, "-v", "0", "-O", "-i", str(STREAMS), "-F", "r", "-w", os.path.join(tmp, "out"), data],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rows = []
    for line in open(os.path.join(tmp, "out", "finalAnalysisReport.txt")):
        m = re.match(r"^\s*((?:\d+\s+){10})\s*([0-9.]+|----)\s+(\*)?\s*(\d+)/(\d+)\s+(\*)?\s*(\w+)\s*$", line)
        if m:
            bins = [int(x) for x in m.group(1).split()]
            uni = float(m.group(2)) if m.group(2) != "----" else -1.0
            rows.append((bins, uni, int(m.group(4)), int(m.group(5)), m.group(7), bool(m.group(3)), bool(m.group(6))))
    assert len(rows) == 188, len(rows)
    np.savez_compressed(os.path.join(HERE, "assess_lcg_64.npz"),
                        streams=np.int64(STREAMS), bins=np.array([r[0] for r in rows], dtype=np.int64),
                        uniformity=np.array([r[1] for r in rows]), passed=np.array([r[2] for r in rows], dtype=np.int64),
                        sample=np.array([r[3] for r in rows], dtype=np.int64),
                        names=np.array([r[4] for r in rows]),
                        uniformity_failed=np.array([r[5] for r in rows]), proportion_failed=np.array([r[6] for r in rows]))
    print("wrote assess_lcg_64.npz: %d columns" % len(rows))


if __name__ == "__main__":
    main()
