#!/usr/bin/env python3
"""Golden vector for -F a input, from the UNMODIFIED reference binary.

The text is the bits of the first 4 LCG streams as '0' / '1' characters with a line break after every 80
digits (tests/golden_util.py:ascii_lines).  The reference seeks to character iteration * n before every
iteration (utilities.c:1682-1683) and then reads n digits skipping white space, so in such a file consecutive
streams overlap; this fixture pins that reading.  Runs oracle/_ref/sts -F a -T 1 -m i -i 3 for jobs 0 and
the p-values go to tests/golden/ascii_lines80_3.npz.  Only runs in the build container (needs oracle/_ref)."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
STS = os.path.join(ROOT, "oracle", "_ref", "sts")
ITER, LINE, SRC_STREAMS = 3, 80, 4


def main():
    from oracle import oracle as O
    from golden_util import ascii_lines
    tmp = tempfile.mkdtemp(prefix="ascii_golden_")
    path = os.path.join(tmp, "ascii.txt")
    ascii_lines(np.unpackbits(O.lcg_bytes(0, SRC_STREAMS)), LINE).tofile(path)
    subprocess.check_call([STS, "-v", "0", "-m", "i", "-T", "1", "-i", str(ITER), "-F", "a", "-w", os.path.join(tmp, "w"), path],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    buf = open(os.path.join(tmp, "w", "sts.0000.%d.1048576.pvalues" % ITER), "rb").read()
    out, off = {}, 0
    while off < len(buf):
        t, cnt = struct.unpack_from("<qq", buf, off)
        off += 16
        out["pv_%d" % t] = np.frombuffer(buf, dtype="<f8", count=cnt, offset=off).reshape(ITER, -1).copy()
        off += 8 * cnt
    np.savez_compressed(os.path.join(HERE, "ascii_lines80_3.npz"), n=np.int64(1 << 20), iterations=np.int64(ITER),
                        line=np.int64(LINE), source_streams=np.int64(SRC_STREAMS), **out)
    print("wrote ascii_lines80_3.npz:", sorted(out))


if __name__ == "__main__":
    main()
