#!/usr/bin/env python3
"""Development aid: wall time of the device generators (1, 2, 4, 5) for B streams of 2^20 bits, and a check of stream
B - 1 against the oracle's restatement for 2, 4, 5 (its own jump, far from the start)."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
with sts_b200.Engine(max_streams=B, test_mask=1 << 1) as eng:
    for gen in (1, 2, 4, 5):
        eng.generate(gen, 0, B)
        eng.download_input(1)
        ts = []
        for r in range(3):
            t0 = time.perf_counter()
            eng.generate(gen, r * B, B)
            eng.download_input(1)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("generator %d: %.3f ms per %d streams (%.1f Gbit/s)" % (gen, min(ts), B, B * (1 << 20) / min(ts) / 1e6))
if len(sys.argv) > 2:
    from oracle import generators as OG
    k = int(sys.argv[2])
    with sts_b200.Engine(max_streams=1, test_mask=1 << 1) as eng:
        for gen in (2, 4, 5):
            eng.generate(gen, k, 1)
            got = hashlib.sha256(eng.download_input(1).tobytes()).hexdigest()
            want = hashlib.sha256(OG.generate(gen, 1 << 20, k + 1)[k]).hexdigest()
            print("generator %d stream %d: %s" % (gen, k, "ok" if got == want else "MISMATCH"))
