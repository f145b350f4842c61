#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: differential fuzzing of the emulated engine (the unmodified kernels on the CPU emulator of
tests/emu) against the oracle: random stream lengths (also not multiples of 32), random valid test parameters, random /
biased / periodic / degenerate inputs; records compared like tests/test_gpu_parity.py::compare_with_oracle.

usage: python tests/emu/fuzz.py [first_seed] [count]      (prints one line per case, exits 1 at the first mismatch)
STS_FUZZ_GPU=1 runs the same cases through the real engine (sts_b200/libstsgpu.so) on cuda:0 instead of the emulator."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_util import emu_binding  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_gpu_parity as G  # noqa: E402


def largest_prime_factor(x):
    p, best = 2, 1
    while p * p <= x:
        while x % p == 0:
            best, x = p, x // p
        p += 1
    return max(best, x) if x > 1 else best


def make_input(rng, n, streams):
    kind = rng.choice(["lcg", "uniform", "biased", "periodic", "sparse", "runs", "const"])
    nb = n // 8
    if kind == "lcg":
        return kind, O.lcg_bytes(int(rng.randint(0, 2000)), streams, n)
    if kind == "uniform":
        return kind, rng.randint(0, 256, size=streams * nb).astype(np.uint8)
    if kind == "biased":
        p = rng.choice([0.02, 0.125, 0.4, 0.6, 0.875, 0.98])
        return "biased%.3f" % p, np.packbits((rng.random_sample(streams * n) < p).astype(np.uint8))
    if kind == "periodic":
        period = int(rng.choice([2, 3, 5, 7, 9, 16, 31, 500, 1032]))
        pat = rng.randint(0, 2, size=period).astype(np.uint8)
        bits = np.resize(pat, streams * n)
        return "periodic%d" % period, np.packbits(bits)
    if kind == "sparse":
        bits = np.zeros(streams * n, dtype=np.uint8)
        bits[rng.randint(0, streams * n, size=int(rng.randint(1, 50)))] = 1
        if rng.randint(0, 2):
            bits ^= 1
        return kind, np.packbits(bits)
    if kind == "runs":
        lens = rng.geometric(1.0 / rng.choice([2, 8, 40, 300]), size=streams * n // 2 + 2)
        bits = np.resize(np.repeat(np.arange(lens.size) & 1, lens), streams * n).astype(np.uint8)
        return kind, np.packbits(bits)
    return kind, np.full(streams * nb, rng.choice([0x00, 0xFF, 0x55, 0xF0]), dtype=np.uint8)


def one_case(sts, seed):
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1000, 1024, 2048, 4104, 38912, 65536, 100008, 387840, 400000, 500000, 904960, 1000000, 1000008,
                        1 << 20, (1 << 20) + 8, 1200000, 1 << 21]))
    streams = int(rng.randint(1, 4))
    prm = dict(n=n,
               block_frequency_M=int(rng.choice([20, 128, 1000, 10000, 16384, 20000, n // 3 + 1])),
               non_overlapping_m=int(rng.choice([8, 9, 9, 10, 11])),
               overlapping_m=9,
               apen_m=int(rng.randint(1, 13)),
               serial_m=int(rng.randint(3, 17)),
               linear_complexity_M=int(rng.choice([500, 501, 511, 512, 777, 1022, 1023, 1500])))
    mask = sts.ALL_TESTS
    if not sts.dft_supported(n) or largest_prime_factor(n // 2) > 3000:      # the oracle's FFT is O(N p) for a prime factor p
        mask &= ~(1 << 7)
    if rng.randint(0, 3) == 0:
        mask &= int(rng.randint(0, 1 << 16)) | (1 << 1)
    kind, raw = make_input(rng, n, streams)
    raw = np.ascontiguousarray(raw[:streams * n // 8])
    with sts.Engine(max_streams=streams, test_mask=mask, pipeline_depth=int(rng.randint(1, 4)), **prm) as eng:
        lay = eng.layout
        want_mask = O.layout(O.make_params(test_mask=mask, **prm)).enabled_mask
        assert lay.enabled_mask == want_mask, "enabled mask %#x, oracle %#x" % (lay.enabled_mask, want_mask)
        pv, st, w = eng.run(raw, streams, int(rng.randint(0, 1000)))
        G.compare_with_oracle(sts, O, eng, mask, prm, raw, streams, pv, st, w)
    return "n=%d streams=%d mask=%#x %s bfM=%d m=%d/%d/%d M=%d" % (n, streams, lay.enabled_mask, kind, prm["block_frequency_M"],
                                                                  prm["non_overlapping_m"], prm["apen_m"], prm["serial_m"],
                                                                  prm["linear_complexity_M"])


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    if os.environ.get("STS_FUZZ_GPU", "") not in ("", "0"):	# the same cases through the product on a B200
        import sts_b200 as sts
    else:
        sts = emu_binding()
    for seed in range(first, first + count):
        try:
            print("seed %d ok: %s" % (seed, one_case(sts, seed)), flush=True)
        except Exception as e:                                 # noqa: BLE001
            print("seed %d FAILED: %s: %s" % (seed, type(e).__name__, str(e)[:1500]), flush=True)
            raise SystemExit(1)


if __name__ == "__main__":
    main()
