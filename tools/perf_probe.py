#!/usr/bin/env python3
"""Development aid: per-kernel-group device times for B streams of on-device LCG data, one batch at a
time (groups serialised, then overlapped), and the throughput with `depth` batches in flight."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = sts_b200.Engine(max_streams=B, pipeline_depth=depth)
for d in range(depth):
    eng.generate_lcg(0, B)
    eng.submit_resident(B, 0)
for d in range(depth):
    eng.wait()
for serial in (1, 0):
    eng.set_profiling(serial)
    acc = {}
    tot = []
    wall = []
    for r in range(reps + 2):
        t0 = time.perf_counter()
        eng.submit_resident(B, 0)
        eng.wait()
        t1 = time.perf_counter()
        if r < 2:
            continue
        wall.append((t1 - t0) * 1e3)
        tot.append(eng.last_batch_ms())
        for name, ms in eng.kernel_times():
            acc.setdefault(name, []).append(ms)
    print("---- %s, %d streams x 2^20 bits" % ("serialised" if serial else "one batch, two streams", B))
    for name, v in acc.items():
        print("  %-44s %8.3f ms" % (name, sum(v) / len(v)))
    bt = sum(tot) / len(tot)
    print("  batch device time %.3f ms, wall %.3f ms -> %.1f Gbit/s (device), %.1f Gbit/s (wall)" % (
        bt, sum(wall) / len(wall), B * 2 ** 20 / bt / 1e6, B * 2 ** 20 / (sum(wall) / len(wall)) / 1e6))
# pipelined throughput
n = max(4 * reps, 8)
pending = 0
for phase in (0, 1):
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
    t1 = time.perf_counter()
print("---- pipelined, %d batches in flight: %.3f ms per batch -> %.1f Gbit/s" % (
    depth, (t1 - t0) * 1e3 / n, n * B * 2 ** 20 / (t1 - t0) / 1e9))
