#!/usr/bin/env python3
"""Development aid: per-kernel-group device times for BASELINE config 4
(n = 10^7, LinearComplexity M = 5000, template m = 10, Serial/ApEn m = 16)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = sts_b200.Engine(max_streams=B, n=10 ** 7, linear_complexity_M=5000, non_overlapping_m=10, apen_m=16, serial_m=16)
eng.generate_lcg(0, B)
eng.set_profiling(1)
for r in range(REPS):
    eng.submit_resident(B, 0)
    eng.wait()
tot = eng.last_batch_ms()
for name, ms in eng.kernel_times():
    print("  %-44s %9.3f ms" % (name, ms))
print("  batch %.3f ms for %d streams of 10^7 bits -> %.2f Gbit/s" % (tot, B, B * 1e7 / tot / 1e6))
