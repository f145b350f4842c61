#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: fuzz of the "-F a" packer (stsgpu_submit_ascii, k_ascii.cu) on the emulated engine against the
reference's reading rule (parseBitsASCIIInput, utilities.c:1682-1697: iteration k = the first n digits at or after
character k * n, white space skipped; pinned to the reference binary by tests/golden/ascii_lines80_3.npz): random
white-space runs of random kinds and lengths, texts that end early, streams that start inside white space.
usage: python tests/emu/fuzz_ascii.py first_seed last_seed"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_util import emu_binding  # noqa: E402
from golden_util import ascii_seek_rule_streams  # noqa: E402

if os.environ.get("STS_FUZZ_GPU", "") not in ("", "0"):	# the same cases through the product on a B200
    import sts_b200 as sts
else:
    sts = emu_binding()
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1024, 4096, 8200, 65536]))
    streams = int(rng.randint(1, 6))
    ndig = int(n * (streams + rng.choice([-0.5, 0.0, 0.3, 1.0])))
    digits = rng.randint(0, 2, size=max(ndig, 1)).astype(np.uint8) + ord("0")
    # white space after a digit with probability p, in runs of geometric length
    p = float(rng.choice([0.0, 0.001, 0.02, 0.3]))
    gaps = (rng.random_sample(digits.size) < p) * rng.geometric(float(rng.choice([0.9, 0.3, 0.02])), size=digits.size)
    lead = int(rng.choice([0, 0, 1, 70, n // 2]))
    out = [np.full(lead, ord(" "), dtype=np.uint8)]
    ws = np.array([ord(" "), ord("\n"), ord("\t"), ord("\r")], dtype=np.uint8)
    pos = np.concatenate([[0], np.cumsum(1 + gaps)])
    text = np.full(int(pos[-1]), ord(" "), dtype=np.uint8)
    text[:] = rng.choice(ws, size=text.size)
    text[pos[:-1]] = digits
    text = np.concatenate([out[0], text])[:(streams + 1) * n]
    want, complete = ascii_seek_rule_streams(text, n, streams)
    with sts.Engine(max_streams=streams, n=n, test_mask=(1 << 1) | (1 << 4)) as eng:
        eng.submit_ascii(text, streams, 0)
        eng.wait()
        done, badoff = eng.ascii_status()
        got = eng.download_input(streams).reshape(streams, n // 8)
        ok = done == complete and badoff == -1 and (got[:complete] == want[:complete]).all()
        if ok and complete:
            pv_a = eng.results()[0]
            eng.submit(want[:complete].reshape(-1).copy(), complete, 0)
            eng.wait()
            ok = (eng.results()[0] == pv_a).all()
    print("seed", seed, "n", n, "streams", streams, "complete", complete, "ws p", p, "lead", lead, "OK" if ok else "MISMATCH (done %d bad %d)" % (done, badoff),
          flush=True)
    bad += not ok
print("bad", bad)
sys.exit(1 if bad else 0)
