"""CPU: the arithmetic behind sts_b200/csrc/rne_sum.cuh, checked with Python integers against IEEE double addition.

The Universal and ApEn statistics are sequential double sums whose rounding depends on the order
(universal.c:369, approximateEntropy.c:396-407); the kernels reproduce them in parallel.  The claims are:
  (1) inside one binade, adding x >= 0 to a sum with integer significand S is S -> S + (S even ? a0 : a0 + dl)
      with (a0, dl) depending on x only (rne_of_term);
  (2) such maps compose into a map of the same form (rne_compose), so any bracketing - four per lane, then a
      butterfly over the lanes - gives the sequential result;
  (3) for fixed-point terms X = x 2^58 without an exact tie the whole group is the integer sum of (X + half) >> sh
      (warp_rne_sum_rows_fixed).
Each is exercised on random terms and on constructed ties / near-ties / binade edges, where the fall-back
conditions (tie seen, sum leaves the binade, term too large) must be the only cases that differ."""
import random
import struct

import pytest


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def from_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def rne_of_term(x, k):
    xb = bits(x)
    if xb == 0:
        return (0, 0)
    e = (xb >> 52) - 1023
    M = (xb & ((1 << 52) - 1)) | (1 << 52)
    sh = min(k - e, 62)
    q, r, half = M >> sh, M & ((1 << sh) - 1), 1 << (sh - 1)
    dl = 0
    if r > half:
        q += 1
    elif r == half:
        dl = -1 if q & 1 else 1
        q += q & 1
    return (q, dl)


def compose(f, g):
    f1, g1 = f[0] + f[1], g[0] + g[1]
    h0 = f[0] + (g1 if f[0] & 1 else g[0])
    h1 = f1 + (g[0] if f1 & 1 else g1)
    return (h0, h1 - h0)


def apply_map(sum_, f):
    sb = bits(sum_)
    k = (sb >> 52) - 1023
    S = (sb & ((1 << 52) - 1)) | (1 << 52)
    Snew = S + (f[0] + f[1] if S & 1 else f[0])
    if Snew >= 1 << 53:
        return None
    return from_bits(((k + 1023) << 52) | (Snew & ((1 << 52) - 1)))


def tree(maps):
    while len(maps) > 1:
        maps = [compose(maps[i], maps[i + 1]) if i + 1 < len(maps) else maps[i] for i in range(0, len(maps), 2)]
    return maps[0]


def sequential(sum_, terms):
    for t in terms:
        sum_ = sum_ + t                       # Python floats are IEEE doubles, round to nearest even
    return sum_


def term_sets(rng, k):
    """terms at least two binades below 2^k: random ones, exact ties and their neighbours in units of 2^(k-52)"""
    u = 2.0 ** (k - 52)
    yield [rng.uniform(1.0, 31.0) for _ in range(128)]
    yield [rng.choice([0.0, 1.0, 2.0, 6.5, 31.999]) for _ in range(128)]
    yield [(rng.randrange(1, 1 << 20) + 0.5) * u for _ in range(128)]                      # every term is a tie
    yield [(rng.randrange(1, 1 << 20) + rng.choice([0.5, 0.5 - 2 ** -20, 0.5 + 2 ** -20, 0.0])) * u for _ in range(128)]
    yield [0.5 * u] * 128                                                                   # alternating parity
    yield [1.5 * u, 0.5 * u] * 64


@pytest.mark.parametrize("k", [3, 10, 19])
def test_maps_reproduce_sequential_addition_under_any_bracketing(k):
    rng = random.Random(k)
    for trial in range(40):
        for terms in term_sets(rng, k):
            terms = [t for t in terms if t < 2.0 ** (k - 1)]
            start = rng.uniform(2.0 ** k, 2.0 ** (k + 1) * 0.98)
            if trial % 3 == 0:
                start = from_bits(bits(start) | 1)          # odd significand
            want = sequential(start, terms)
            maps = [rne_of_term(t, k) for t in terms]
            seq_map = maps[0]
            for m in maps[1:]:
                seq_map = compose(seq_map, m)
            assert seq_map == tree(maps)                      # claim (2): associativity of the composition
            got = apply_map(start, seq_map)
            if want >= 2.0 ** (k + 1):
                assert got is None                            # leaves the binade: reported back, never a wrong value
            else:
                assert got == want                            # claims (1) + (2)


def test_fixed_point_fast_path_is_exact_away_from_ties():
    FX = 58
    rng = random.Random(58)
    import math
    table = [0.0] + [math.log(d) / math.log(2.0) for d in range(1, 5000)]
    for trial in range(300):
        k = rng.choice([6, 12, 19])
        start = rng.uniform(2.0 ** k, 2.0 ** (k + 1) * 0.9)
        terms = [table[rng.randrange(1, 5000)] for _ in range(1024)]
        if trial % 10 == 0:                                   # plant an exact tie: half a unit of the sum's last place
            terms[rng.randrange(1024)] = 3.0 + 2.0 ** (k - 53)
        X = [int(t * 2 ** FX) for t in terms]
        assert all(float(x) / 2 ** FX == t for x, t in zip(X, terms))        # the table is exact in fixed point
        sh = k - 52 + FX
        half, mask = 1 << (sh - 1), (1 << sh) - 1
        tie = any((x & mask) == half for x in X)
        S = (bits(start) & ((1 << 52) - 1)) | (1 << 52)
        Snew = S + sum((x + half) >> sh for x in X)
        want = sequential(start, terms)
        if tie or Snew >= 1 << 53:
            continue                                          # the kernel falls back to the composable maps
        assert from_bits(((k + 1023) << 52) | (Snew & ((1 << 52) - 1))) == want
