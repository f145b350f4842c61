#!/usr/bin/env python3
"""Development aid: marginal cost of each kernel group in the pipelined steady state - the ms per batch
saved when its tests are switched off (two batches in flight, 1024 streams)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

B, depth, n = 1024, 2, 24
groups = {"all": [], "dft": [7], "linear_complexity": [15], "pattern": [11, 14], "universal": [10],
          "scan": [1, 3, 4, 12, 13], "blocks": [2, 5, 9], "rank": [6], "nonoverlapping": [8]}


def run(mask):
    eng = sts_b200.Engine(max_streams=B, pipeline_depth=depth, test_mask=mask)
    for d in range(depth):
        eng.generate_lcg(0, B)
        eng.submit_resident(B, 0)
    for d in range(depth):
        eng.wait()
    best = 1e9
    for rep in range(2):
        pending = 0
        t0 = time.perf_counter()
        for i in range(n):
            eng.submit_resident(B, 0)
            pending += 1
            if pending == depth:
                eng.wait()
                pending -= 1
        while pending:
            eng.wait()
            pending -= 1
        best = min(best, (time.perf_counter() - t0) * 1e3 / n)
    eng.close()
    return best


base = None
for name, off in groups.items():
    mask = sts_b200.ALL_TESTS
    for t in off:
        mask &= ~(1 << t)
    ms = run(mask)
    if base is None:
        base = ms
    print("without %-18s %7.3f ms per batch   marginal cost %6.3f ms" % (name, ms, base - ms))
