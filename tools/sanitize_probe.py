#!/usr/bin/env python3
"""Development aid: one small batch through every kernel family, meant to be run under
`compute-sanitizer --tool memcheck` (or racecheck): default parameters at n = 2^20 (fast DFT, thread-per-block
LinearComplexity), n = 10^6 (radix-10 DFT), config-4 parameters at n = 10^7 (generic DFT, warp-per-block
LinearComplexity, 16-bit pattern histograms with the sliced fallback), a ragged short length, the ASCII packer,
the LCG generator and the assessment tallies."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sts_b200  # noqa: E402


def run(tag, streams, **prm):
    eng = sts_b200.Engine(max_streams=streams, **prm)
    eng.assess_begin(4)
    eng.generate_lcg(0, streams)
    eng.submit_resident(streams, 0)
    eng.wait()
    pv, st, w = eng.results()
    m = eng.assess_finish()
    print(tag, "ok", pv.shape, float(np.nan_to_num(pv[pv > -1]).sum()), int(m["sample"].sum()))
    eng.close()


run("default 2^20", 3)
run("n=10^6", 2, n=10 ** 6)
run("config4 10^7", 1, n=10 ** 7, linear_complexity_M=5000, non_overlapping_m=10, apen_m=16, serial_m=16)
run("ragged 400000", 2, n=400000, test_mask=sts_b200.ALL_TESTS & ~((1 << 9) | (1 << 12) | (1 << 13) | (1 << 15)))

# degenerate streams: all ones / alternating (pattern histogram overflow -> sliced fallback)
n = 1 << 20
eng = sts_b200.Engine(max_streams=2, n=n)
raw = np.concatenate([np.full(n // 8, 0xFF, np.uint8), np.full(n // 8, 0x55, np.uint8)])
eng.submit(raw, 2, 0)
eng.wait()
print("degenerate ok", eng.results()[0].shape)
eng.close()

# ASCII packer with line breaks, short text
n = 1 << 16
bits = (np.arange(n * 3 + 77) * 2654435761 >> 7) & 1
ch = (bits + 48).astype(np.uint8)
text = np.insert(ch, np.arange(61, ch.size, 61), 10)
eng = sts_b200.Engine(max_streams=4, n=n, test_mask=(1 << 1) | (1 << 4))
eng.submit_ascii(text, 4, 0)
eng.wait()
print("ascii ok", eng.ascii_status())
eng.close()

# the device generators (k_gen.cu): ragged lengths, a stream far from the start, every template width of the strided
# non-overlapping count (k_pattern.cu)
for n in (1 << 20, 1000 * 8, 512 * 3 + 8):
    eng = sts_b200.Engine(max_streams=3, n=n, test_mask=1 << 1)
    for gen in (2, 4, 5):
        eng.generate(gen, 7, 3)
        eng.download_input(3)
    eng.close()
print("generators ok")
for m in (8, 10, 12, 13, 15):
    run("templates m=%d" % m, 2, n=(1 << 19) - 40, non_overlapping_m=m, test_mask=1 << 8)
