#!/usr/bin/env python3
"""Development aid: run every golden case through the engine on cuda:0 and print, per case and per
test, whether integer statistics and p-values agree with the golden vectors and with the oracle.
Writes gpurun_out/gpu_check.json.  (The judged tests are tests/test_gpu_*.py; this prints more.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sts_b200  # noqa: E402
from golden_util import golden_cases, golden_input, golden_params, load_golden, rel_err  # noqa: E402
from oracle import oracle as O  # noqa: E402


def compare(name, pv, st, w, lay, opv, ost, ow, olay, report):
    ok_all = True
    for t in range(1, 16):
        if not (lay.enabled_mask >> t) & 1:
            continue
        a = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
        b = opv[:, olay.pv_off[t]:olay.pv_off[t] + olay.pv_cnt[t]]
        r = rel_err(a, b)
        ad = np.abs(a - b)
        sa = st[:, lay.st_off[t]:lay.st_off[t] + lay.st_cnt[t]]
        sb = ost[:, olay.st_off[t]:olay.st_off[t] + olay.st_cnt[t]]
        st_ok = bool((sa == sb).all())
        if t in (10, 11) and not st_ok:
            # double bit patterns: report as relative error instead
            da = sa[:, :2 if t == 11 else 1].copy().view(np.float64)
            db = sb[:, :2 if t == 11 else 1].copy().view(np.float64)
            st_note = "fp-stat rel %.3g; int part %s" % (rel_err(da, db).max(), bool((sa[:, 2:] == sb[:, 2:]).all()) if t == 11 else "-")
        else:
            st_note = ""
        w_ok = True
        if t == 8:
            w_ok = bool((w == ow).all())
        line = dict(case=name, test=t, name=sts_b200.TEST_NAMES[t], pv_max_rel=float(r.max()), pv_max_abs=float(ad.max()),
                    stats_equal=st_ok, w_equal=w_ok, note=st_note)
        report.append(line)
        good = r.max() <= 1e-9 and st_ok and w_ok
        ok_all &= good
        flag = "ok " if good else "BAD"
        print("  %s %2d %-24s pv rel %.3g abs %.3g  stats %s %s %s" % (flag, t, sts_b200.TEST_NAMES[t], r.max(), ad.max(),
                                                                   st_ok, "" if t != 8 else "W %s" % w_ok, st_note))
        if not st_ok and t not in (10, 11):
            bad = np.argwhere(sa != sb)[:6]
            for (i, j) in bad:
                print("       stream %d slot %d: gpu %d oracle %d" % (i, j, sa[i, j], sb[i, j]))
    return ok_all


def main():
    cases = sys.argv[1:] or [c for c in golden_cases()]
    report = []
    all_ok = True
    for name in cases:
        g = load_golden(name)
        prm = golden_params(g)
        streams = int(g["streams"])
        raw = golden_input(g, O)
        n = prm["n"]
        mask = sts_b200.ALL_TESTS
        note = ""
        if n & (n - 1):
            mask &= ~(1 << 7)
            note = " (DFT masked: n not a power of two)"
        print("== case %s: n=%d streams=%d%s" % (name, n, streams, note))
        t0 = time.time()
        try:
            eng = sts_b200.Engine(max_streams=streams, test_mask=mask, **prm)
        except sts_b200.EngineError as e:
            print("  engine create failed:", e)
            all_ok = False
            continue
        pv, st, w = eng.run(raw, streams)
        lay = eng.layout
        print("  engine: %.3f s, batch %.3f ms device time, launches %d" % (time.time() - t0, eng.last_batch_ms(), eng.launch_count()))
        oprm = O.make_params(test_mask=mask, **prm)
        opv, ost, ow, olay = O.run(oprm, raw, streams, nthreads=os.cpu_count())
        if lay.enabled_mask != olay.enabled_mask:
            print("  enabled mask differs: %#x vs %#x" % (lay.enabled_mask, olay.enabled_mask))
            all_ok = False
            continue
        all_ok &= compare(name, pv, st, w, lay, opv, ost, ow, olay, report)
        # golden p-values (reference binary)
        for t in range(1, 16):
            if (lay.enabled_mask >> t) & 1 and ("pv_%d" % t) in g.files:
                a = pv[:, lay.pv_off[t]:lay.pv_off[t] + lay.pv_cnt[t]]
                r = rel_err(a, g["pv_%d" % t])
                if r.max() > 1e-9:
                    print("  golden: test %d p-value rel err %.3g" % (t, r.max()))
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("ALL OK" if all_ok else "SOME MISMATCHES")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
