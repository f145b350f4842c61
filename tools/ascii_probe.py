#!/usr/bin/env python3
"""Development aid: throughput of the -F a path (text upload + device packing + all tests), B streams per batch from
pinned host memory, three batches in flight; and of the packer alone (frequency test only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = 1 << 20
for tag, mask in (("all 15 tests", sts_b200.ALL_TESTS), ("frequency only", 1 << 1)):
    eng = sts_b200.Engine(max_streams=B, pipeline_depth=3, test_mask=mask)
    bufs = []
    rng = np.random.RandomState(1)
    for d in range(3):
        pb = sts_b200.binding.PinnedBuffer((B + 1) * n)
        pb.array[:] = rng.randint(0, 2, size=(B + 1) * n).astype(np.uint8) + 48
        bufs.append(pb)
    reps, pending = 12, 0
    for phase in (0, 1):
        t0 = time.perf_counter()
        for i in range(reps):
            eng.submit_ascii(bufs[i % 3].array, B, i * B)
            pending += 1
            if pending == 3:
                eng.wait()
                pending -= 1
        while pending:
            eng.wait()
            pending -= 1
        dt = (time.perf_counter() - t0) / reps
    print("%-15s %4d streams/batch: %.2f ms per batch, %.1f GB/s of text, %.1f Gbit/s tested" % (
        tag, B, dt * 1e3, (B + 1) * n / dt / 1e9, B * n / dt / 1e9))
    eng.close()
    for pb in bufs:
        pb.free()
