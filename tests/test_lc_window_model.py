"""CPU: the live-window rule of the LinearComplexity kernels (sts_b200/csrc/k_lc.cu), checked as arithmetic.

Both kernels run Berlekamp-Massey in "sequence-aligned reversed" form (index k <-> position k - 1, the connection
polynomial reversed and aligned under the sequence, shifted by one index per step) and touch only the indices
between a lower bound derived at the start of a phase, min(N0 - L + 1, L), and the top index of the phase:
  - per-thread kernel (M <= 511): phases of 32 steps, whole 32-bit words [lo_idx >> 5, P + 1];
  - warp-per-block kernel (M > 1022): phases ending where the top index reaches a multiple of 256, whole register
    slots of 32 words = 1024 indices (the cyclic word layout).
The model below applies exactly those rules with Python integers and must return the linear complexity of the
textbook algorithm (linearComplexity.c:283-340) - in particular on the inputs where L stays small, jumps late or
never moves, which random data never produces."""
import random

import pytest


def bm_textbook(bits):
    M = len(bits)
    c, b = [0] * (M + 2), [0] * (M + 2)
    c[0] = b[0] = 1
    L, m = 0, -1
    for N in range(M):
        d = bits[N]
        for i in range(1, L + 1):
            d ^= c[i] & bits[N - i]
        if d:
            t = c[:]
            for j in range(0, M - N + m + 1):
                if b[j] and j + N - m < len(c):
                    c[j + N - m] ^= 1
            if L <= N // 2:
                L, m, b = N + 1 - L, N, t
    return L


def bm_windowed(bits, phase_end, unit, units):
    """phase_end(Na) -> Nb; the live index range of a phase is [unit * (lo_idx // unit), unit * (top_unit + 1))"""
    M = len(bits)
    S = 0
    for k, bit in enumerate(bits):
        if bit:
            S |= 1 << (k + 1)
    C, B, L, Na = 1 << 1, 1 << 0, 0, 0
    while Na < M:
        Nb = min(M, phase_end(Na))
        lo = unit * (min(Na - L + 1, L) // unit)
        hi = unit * (min((Nb + 1) // unit, units - 1) + 1)
        mask = ((1 << hi) - 1) ^ ((1 << lo) - 1)
        for N in range(Na, Nb):
            if bin(C & S & mask).count("1") & 1:
                if 2 * L <= N:
                    C, B = (C & ~mask) | ((C ^ B) & mask), (B & ~mask) | (C & mask)
                    L = N + 1 - L
                else:
                    C = (C & ~mask) | ((C ^ B) & mask)
            C = (C & ~mask) | (((C & mask) << 1) & mask)      # advance to N + 1; nothing enters from below the window
        Na = Nb
    return L


def inputs(M, rng):
    yield "zeros", [0] * M
    yield "ones", [1] * M
    yield "0101", [i & 1 for i in range(M)]
    yield "0011", [(i >> 1) & 1 for i in range(M)]
    for name, pos in (("one at the start", 0), ("one in the middle", M // 2), ("one at the end", M - 1)):
        v = [0] * M
        v[pos] = 1
        yield name, v
    yield "biased", [1 if rng.random() < 0.875 else 0 for _ in range(M)]
    for r in range(2):
        yield "random %d" % r, [rng.getrandbits(1) for _ in range(M)]
    lfsr = [rng.getrandbits(1) for _ in range(40)]
    while len(lfsr) < M:
        lfsr.append(lfsr[-7] ^ lfsr[-40])
    yield "lfsr of length 40", lfsr
    for name, pos in (("flip near the end", M - 3), ("flip past the middle", M // 2 + 17)):
        v = lfsr[:]
        v[pos] ^= 1
        yield "lfsr, " + name, v


@pytest.mark.parametrize("M", [500, 511, 64])
def test_per_thread_word_window(M):
    """lc_phases / lc_dispatch: 32-step phases, top word P + 1, bottom word (min(N0 - L + 1, L)) >> 5"""
    W = (M + 1 + 31) // 32
    rng = random.Random(M)
    for name, bits in inputs(M, rng):
        # the kernel's top word is P + 1 for N in [32 P, 32 P + 32): express it through the same (Nb + 1) // unit rule
        got = bm_windowed(bits, lambda Na: (Na // 32 + 1) * 32, 32, W)
        assert got == bm_textbook(bits), name


@pytest.mark.parametrize("M", [1100, 2048, 5000])
def test_warp_slot_window(M):
    """lc_kernel_warp: phases end at N + 2 = multiple of 256, live slots of 1024 indices"""
    wpl = ((M + 1 + 31) // 32 + 31) // 32
    rng = random.Random(M)
    for name, bits in inputs(M, rng):
        got = bm_windowed(bits, lambda Na: ((Na + 2) // 256 + 1) * 256 - 2, 1024, wpl)
        assert got == bm_textbook(bits), name
