#!/usr/bin/env python3
"""Development aid: end-to-end wall time of the drop-in (integration/_ref/sts_gpu, the reference's own main()) and of
the mirror host driver (sts_b200/host/stsb200) on a file of LCG streams in /dev/shm, next to the reference binary on a
small sample.  usage: dropin_probe.py [streams=32768]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
path = "/dev/shm/lcg_%d.bin" % streams
B = 1024
with sts_b200.Engine(max_streams=B, test_mask=2) as eng, open(path, "wb") as f:
    for s0 in range(0, streams, B):
        eng.generate_lcg(s0, B)
        eng.download_input(B).tofile(f)
print("file: %s, %.2f GiB" % (path, os.path.getsize(path) / 2.0 ** 30))
for name, exe, extra in (("sts_gpu (drop-in: reference main + glue.c)", os.path.join(ROOT, "integration", "_ref", "sts_gpu"), []),
                         ("stsb200 (mirror host driver)", os.path.join(ROOT, "sts_b200", "host", "stsb200"), [])):
    for mode in ("i", "b"):
        w = "/dev/shm/probe_w"
        subprocess.run(["rm", "-rf", w])
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-v", "1", "-m", mode, "-i", str(streams), "-F", "r", "-w", w] + extra + [path],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.perf_counter() - t0
        print("%-46s -m %s: rc %d, %.2f s wall -> %.1f Gbit/s end to end (process start, engine creation, file read, tests, "
              "record keeping, %s)" % (name, mode, r.returncode, dt, streams * (1 << 20) / dt / 1e9,
                                       ".pvalues written" if mode == "i" else "assessment, result.txt"))
        for line in r.stdout.splitlines():
            if "iterate phase" in line:
                print("    " + line.strip())
subprocess.run(["rm", "-rf", "/dev/shm/probe_w", path])
