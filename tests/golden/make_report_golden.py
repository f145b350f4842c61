#!/usr/bin/env python3
"""Golden report bodies from the UNMODIFIED reference binary, so that the result.txt tests also run where
oracle/_ref is absent:
  result_lcg32_body.txt    `sts -i 32 -F r` (mode b, sqrt(32) = 5 uniformity bins) on the first 32 LCG streams
  assess_lcg8_body.txt     `sts -m i -i 8` on the first 8 LCG streams, then `sts -m a -d DIR -i 8`
Each file holds result.txt from its first "- - - -" rule on (the lines above name the data source and directory).
Only runs in the build container (needs oracle/_ref)."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
STS = os.path.join(ROOT, "oracle", "_ref", "sts")


def body(path):
    txt = open(path).read()
    return txt[txt.index("- - - -"):]


def run(*args):
    subprocess.check_call([STS, "-v", "0"] + list(args), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def main():
    from oracle import oracle as O
    tmp = tempfile.mkdtemp(prefix="report_golden_")
    p32, p8 = os.path.join(tmp, "lcg32.bin"), os.path.join(tmp, "lcg8.bin")
    O.lcg_bytes(0, 32).tofile(p32)
    O.lcg_bytes(0, 8).tofile(p8)
    run("-i", "32", "-F", "r", "-w", os.path.join(tmp, "b"), p32)
    open(os.path.join(HERE, "result_lcg32_body.txt"), "w").write(body(os.path.join(tmp, "b", "result.txt")))
    run("-m", "i", "-i", "8", "-F", "r", "-w", os.path.join(tmp, "pv"), p8)
    run("-m", "a", "-d", os.path.join(tmp, "pv"), "-i", "8", "-w", os.path.join(tmp, "a"))
    open(os.path.join(HERE, "assess_lcg8_body.txt"), "w").write(body(os.path.join(tmp, "a", "result.txt")))
    print("wrote result_lcg32_body.txt, assess_lcg8_body.txt")


if __name__ == "__main__":
    main()
