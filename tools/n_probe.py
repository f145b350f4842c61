#!/usr/bin/env python3
"""Development aid: per-kernel-group device times at another stream length (default parameters).
usage: n_probe.py <n> [streams]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

n = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mask = sts_b200.ALL_TESTS if sts_b200.dft_supported(n) else sts_b200.ALL_TESTS & ~(1 << 7)
eng = sts_b200.Engine(max_streams=B, n=n, test_mask=mask)
eng.generate_lcg(0, B)
eng.set_profiling(1)
for r in range(3):
    eng.submit_resident(B, 0)
    eng.wait()
tot = eng.last_batch_ms()
for name, ms in eng.kernel_times():
    print("  %-44s %9.3f ms" % (name, ms))
print("  batch %.3f ms for %d streams of %d bits -> %.2f Gbit/s (groups serialised)" % (tot, B, n, B * n / tot / 1e6))
