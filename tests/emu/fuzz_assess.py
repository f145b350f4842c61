#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: fuzz of the device-side assessment tallies (stsgpu_assess_begin/add/finish) on the emulated engine
against the oracle restatement of X_metrics (frequency.c:631-894): random p-value columns salted with edge values - bin
boundaries, alpha and its neighbours, 0, 1, -0.0, denormal-small, NON_P_VALUE - streamed in two pieces per test.
usage: python tests/emu/fuzz_assess.py first_seed last_seed"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from emu_util import emu_binding
from oracle import oracle as O
from golden_util import rel_err
if os.environ.get("STS_FUZZ_GPU", "") not in ("", "0"):	# the same cases through the product on a B200
    import sts_b200 as sts
else:
    sts = emu_binding()
NONP = -99999999.0
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.RandomState(seed)
    mask = int(rng.choice([0xFFFE, (1 << 1) | (1 << 12) | (1 << 13), (1 << 8) | (1 << 3), 0x3006]))
    prm = dict(n=1 << 20, non_overlapping_m=int(rng.choice([8, 9])))
    lay_o = O.layout(O.make_params(test_mask=mask, **prm))
    iters = int(rng.randint(1, 300))
    bins = int(rng.choice([1, 2, 5, 10, 17, int(np.sqrt(iters)) or 1]))
    alpha = float(rng.choice([0.01, 0.001, 0.05]))
    pv = rng.random_sample((iters, lay_o.npv))
    # edge values: bin boundaries, alpha itself, 0, 1, values outside [0, 1], the "not possible" marker
    edge = np.array([0.0, 1.0, alpha, np.nextafter(alpha, 0), np.nextafter(alpha, 1), 1e-300, 1.0 - 2 ** -53, NONP, -0.0] +
                    [k / bins for k in range(bins + 1)])
    sel = rng.random_sample(pv.shape) < 0.15
    pv[sel] = rng.choice(edge, size=int(sel.sum()))
    with sts.Engine(max_streams=1, test_mask=mask, **prm) as eng:
        lay = eng.layout
        eng.assess_begin(bins, alpha, 0.0001)
        for t in range(1, 16):
            if (lay.enabled_mask >> t) & 1:
                col = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]].reshape(-1)
                cut = int(rng.randint(0, col.size + 1))              # streamed in two pieces, like a file through a buffer
                eng.assess_add(t, col[:cut], 0)
                if cut < col.size:
                    eng.assess_add(t, col[cut:], cut)
        a = eng.assess_finish()
    o = O.assess(lay_o, pv, bins, alpha, 0.0001)
    ok = all((a[k] == o[k]).all() for k in ("sample", "toolow", "passed", "freq", "flags"))
    with np.errstate(all="ignore"):
        ok = ok and all(rel_err(a[k], o[k]).max() <= 1e-9 or (np.isnan(a[k]) == np.isnan(o[k])).all() and rel_err(np.nan_to_num(a[k]), np.nan_to_num(o[k])).max() <= 1e-9 for k in ("tmin", "tmax", "chi2", "uniformity"))
    print("seed", seed, "iters", iters, "bins", bins, "alpha", alpha, "mask %#x" % mask, "OK" if ok else "MISMATCH", flush=True)
    if not ok:
        for k in ("sample", "toolow", "passed", "freq", "flags", "tmin", "tmax", "chi2", "uniformity"):
            d = np.argwhere(np.asarray(a[k]) != np.asarray(o[k]))
            if d.size:
                print("   ", k, d[:5].tolist(), np.asarray(a[k])[tuple(d[0])], np.asarray(o[k])[tuple(d[0])])
        bad += 1
print("bad", bad)
sys.exit(1 if bad else 0)
