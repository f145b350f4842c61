#!/usr/bin/env python3
"""Development aid: do two kernel groups overlap?  Runs a batch with only the given tests enabled,
serialised and overlapped.  usage: pair_probe.py <mask-as-test-numbers, e.g. 7,15> [streams]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

tests = [int(x) for x in sys.argv[1].split(",")]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mask = 0
for t in tests:
    mask |= 1 << t
eng = sts_b200.Engine(max_streams=B, test_mask=mask)
eng.generate_lcg(0, B)
for serial in (1, 0):
    eng.set_profiling(serial)
    tot = []
    kt = {}
    for r in range(5):
        eng.submit_resident(B, 0)
        eng.wait()
        if r >= 2:
            tot.append(eng.last_batch_ms())
            for name, ms in eng.kernel_times():
                kt.setdefault(name, []).append(ms)
    print("tests %s %-10s batch %.3f ms  %s" % (tests, "serial" if serial else "overlap", sum(tot) / len(tot),
                                             {k: round(sum(v) / len(v), 3) for k, v in kt.items() if sum(v) / len(v) > 0.01}))
