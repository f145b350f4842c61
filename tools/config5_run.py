#!/usr/bin/env python3
"""BASELINE config 5 on one multi-GPU box: the reference's distributed flow (README.md:129-213) at the scale where the
reference itself cannot be assessed in memory.  One `stsb200 -g 1 -m i -D k -j k -i I` per GPU tests its `-j` slice of
the linear congruential generator's stream sequence (generated on the device by skip-ahead: no input file) and writes
DIR/sts.000k.I.1048576.pvalues; then ONE `stsb200 -m a -d DIR` streams every file through the GPU tallies
(stsgpu_assess_add, a fixed 8 MiB buffer) and writes result.txt.  Wall times and the peak resident set of both phases
are recorded: `sts -m a` keeps every p-value plus a 64-byte struct per template p-value (defs.h:346-352) in memory,
9.5 KB per stream, 79 GB at 2^23 streams.

usage: config5_run.py <gpus> <iterations per GPU> <workdir> [--keep]"""
import json
import os
import resource
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sts_b200", "host", "stsb200")
G, I, DIR = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.makedirs(DIR, exist_ok=True)


def peak_rss_of(cmd, **kw):
    """run to completion; (seconds, peak RSS in MiB of that child alone) via wait4"""
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, **kw)
    _, status, ru = os.wait4(p.pid, 0)
    p.returncode = os.waitstatus_to_exitcode(status)
    assert p.returncode == 0, (cmd, p.returncode)
    return time.perf_counter() - t0, ru.ru_maxrss / 1024.0


t0 = time.perf_counter()
procs = []
for k in range(G):
    procs.append(subprocess.Popen([EXE, "-v", "1", "-g", "1", "-m", "i", "-D", str(k), "-j", str(k), "-i", str(I), "-w", DIR],
                                  stdout=open(os.path.join(DIR, "iterate_%d.log" % k), "w"), stderr=subprocess.STDOUT))
rss_i = []
for p in procs:
    _, status, ru = os.wait4(p.pid, 0)
    assert os.waitstatus_to_exitcode(status) == 0, open(os.path.join(DIR, "iterate_%d.log" % procs.index(p))).read()[-2000:]
    rss_i.append(ru.ru_maxrss / 1024.0)
t_iter = time.perf_counter() - t0
files = sorted(f for f in os.listdir(DIR) if f.endswith(".pvalues"))
size = sum(os.path.getsize(os.path.join(DIR, f)) for f in files)
t_assess, rss_a = peak_rss_of([EXE, "-v", "1", "-m", "a", "-d", DIR, "-i", str(G * I), "-w", os.path.join(DIR, "assess")],
                              stdout=open(os.path.join(DIR, "assess.log"), "w"), stderr=subprocess.STDOUT)
result = open(os.path.join(DIR, "assess", "result.txt")).read()
n = 1 << 20
out = {"config": "BASELINE config 5: %d GPUs x %d streams of 2^20 bits (%.2f Tbit in total), -m i per GPU, then -m a" % (G, I, G * I * n / 1e12),
       "gpus": G, "streams_per_gpu": I, "streams": G * I, "bits_tested": G * I * n,
       "iterate_seconds": t_iter, "iterate_gbit_per_s": G * I * n / t_iter / 1e9, "iterate_peak_rss_mib_per_process": max(rss_i),
       "pvalues_files": len(files), "pvalues_bytes": size,
       "assess_seconds": t_assess, "assess_peak_rss_mib": rss_a,
       "reference_assess_memory_estimate_gib": G * I * (188 * 8 + 148 * 64) / 2.0 ** 30,
       "iterate_log_rank0": open(os.path.join(DIR, "iterate_0.log")).read()[-600:],
       "result_txt_head": result[:3000]}
print(json.dumps(out))
if "--keep" not in sys.argv:
    shutil.rmtree(DIR, ignore_errors=True)
