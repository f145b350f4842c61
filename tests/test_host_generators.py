"""CPU: the host's chained generators (sts_b200/host/sts_generators.c: 3, 6, 7, 8, 9) against the vectors of the
unmodified reference tool and against the oracle's restatement."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import generators as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "generators.json")))
N = GOLD["n"]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostgen") / "libstsgen.so")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "sts_b200", "host", "sts_generators.c")])
    lib = C.CDLL(so)
    lib.sts_gen_open.restype = C.c_void_p
    lib.sts_gen_open.argtypes = [C.c_int, C.c_long]
    lib.sts_gen_next.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    lib.sts_gen_set_ms_initial.argtypes = [C.c_void_p, C.c_char_p]
    lib.sts_gen_close.argtypes = [C.c_void_p]
    return lib


def make(lib, gen, n, streams, ms_initial=None, calls=1):
    h = lib.sts_gen_open(gen, n)
    assert h
    if ms_initial is not None:
        lib.sts_gen_set_ms_initial(h, ms_initial)
    out = []
    for _ in range(calls):
        buf = np.full(streams * n // 8, 0xAA, dtype=np.uint8)
        assert lib.sts_gen_next(h, buf.ctypes.data, streams) == 0
        out += [buf[k * n // 8:(k + 1) * n // 8].tobytes() for k in range(streams)]
    lib.sts_gen_close(h)
    return out


@pytest.mark.parametrize("gen", [3, 6, 9])
def test_matches_the_reference_tool(lib, gen):
    want = GOLD[str(gen)]["sha256"]
    got = make(lib, gen, N, len(want))
    assert [hashlib.sha256(s).hexdigest() for s in got] == want


def test_micali_schnorr_reproduces_a_recorded_reference_run(lib):
    e = GOLD["8"]
    got = make(lib, 8, N, len(e["sha256"]), ms_initial=bytes.fromhex(e["ms_initial"]))
    assert [hashlib.sha256(s).hexdigest() for s in got] == e["sha256"]
    # and the engine's own definition (zero) is the oracle's with ms_initial = 0
    assert make(lib, 8, 4096, 3) == G.micali_schnorr(4096, 3)


def test_bbs_whole_first_stream_matches_the_reference_tool(lib):
    """2^20 passes of the tool's in-place square: 72 s in the tool, a few seconds here"""
    got = make(lib, 7, N, 1)
    assert hashlib.sha256(got[0]).hexdigest() == GOLD["7"]["sha256"][0]


@pytest.mark.parametrize("gen,n", [(3, 1000), (6, 520), (7, 776), (8, 1600), (8, 840), (9, 168), (9, 160 * 5)])
def test_ragged_stream_lengths_and_repeated_calls_match_the_oracle(lib, gen, n):
    """a stream that ends inside a block drops the rest of the block; the state carries across calls"""
    assert make(lib, gen, n, 3, calls=2) == G.generate(gen, n, 6)


def test_only_the_chained_generators_are_made_here(lib):
    for g in (0, 1, 2, 4, 5, 10):
        assert not lib.sts_gen_open(g, 1024)
    assert not lib.sts_gen_open(3, 1001)


def test_random_lengths_and_counts_match_the_oracle(lib):
    import random
    rnd = random.Random(5)
    for _ in range(25):
        gen = rnd.choice([3, 6, 7, 8, 9])
        n = 8 * rnd.randint(1, 300 if gen in (6, 7) else 3000)
        k = rnd.randint(1, 4)
        assert make(lib, gen, n, k) == G.generate(gen, n, k), (gen, n, k)
