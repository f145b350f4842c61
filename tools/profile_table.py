#!/usr/bin/env python3
"""Turn an `ncu --set full ... --page raw --csv` export of ONE batch (all kernels, serialised) into the
per-kernel roofline table of profiles/: duration, algorithmic GB/s vs the measured HBM peak, DRAM
traffic, and the pipe that bounds the kernel.  usage: profile_table.py raw.csv streams bits [traffic.json] > table.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, B, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
peak = 6549.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


def to_us(r, k):
    v, u = f(r, k), units[ix[k]]
    return v / 1e3 if u.startswith("ns") else (v * 1e3 if u.startswith("ms") else (v * 1e6 if u == "s" else v))


def to_bytes(r, k):
    v, u = f(r, k), units[ix[k]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


agg = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, alu=0.0, fp64=0.0, lsu=0.0, issue=0.0, occ=0.0, regs=0))
    us = to_us(r, "gpu__time_duration.sum")
    a["n"] += 1
    a["us"] += us
    a["rd"] += to_bytes(r, "dram__bytes_read.sum")
    a["wr"] += to_bytes(r, "dram__bytes_write.sum")
    for key, met in (("alu", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     ("fp64", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("lsu", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                     ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        a[key] += f(r, met) * us
    a["regs"] = int(f(r, "launch__registers_per_thread"))

# launches of each kernel in ONE batch of B streams (the capture window may straddle two batches: every kernel is
# scaled to its per-batch launch count, i.e. average launch x launches per batch)
per_batch = {"dft_cols16": max(1, B // 128), "dft_rows16": max(1, B // 128), "dft_rows16_tma": max(1, B // 128)}
for name, a in agg.items():
    want = per_batch.get(name, 1)
    if a["n"] != want:
        k = want / a["n"]
        for key in ("us", "rd", "wr", "alu", "fp64", "lsu", "issue", "occ"):
            a[key] *= k
        a["n"] = want
stream_bytes = n // 8
alg = {  # algorithmic bytes per stream of one launch of the kernel
    "dft_cols16": stream_bytes + 16 * (n // 2), "dft_rows16": 16 * (n // 2), "dft_rows16_tma": 16 * (n // 2),
    "universal_kernel": 113120 if n == 1 << 20 else stream_bytes,
    "universal_dist_kernel": 113120 if n == 1 << 20 else stream_bytes, "universal_sum_kernel": 0,
    "tails_slots_kernel": 0, "tails_cusum_kernel": 0, "tails_apen_kernel": 0, "cyclic_hist_kernel": 0,
}
print("| kernel | launches | total us | algorithmic GB/s | of %.0f GB/s HBM (measured) | DRAM traffic / algorithmic | "
      "INT ALU pipe %% | FP64 pipe %% | LSU wavefront %% | issue %% | warps active %% | regs |" % peak)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    us = a["us"]
    tot += us
    ab = alg.get(name, stream_bytes) * B
    gbs = ab / (us * 1e-6) / 1e9 if us else 0
    traf = (a["rd"] + a["wr"]) / ab if ab else float("nan")
    print("| %s | %d | %.1f | %.0f | %.3f | %.2f | %.0f | %.0f | %.0f | %.0f | %.0f | %d |" % (
        name, a["n"], us, gbs, gbs / peak, traf, a["alu"] / us, a["fp64"] / us, a["lsu"] / us, a["issue"] / us,
        a["occ"] / us, a["regs"]))
if len(sys.argv) > 4:
    # the summary bench.py quotes (roofline.traffic, roofline_all[].frac of the integer / shared-memory pipes): keyed by
    # bench.py's kernel group names and stamped with the hash of the CUDA sources the capture was taken on
    sys.path.insert(0, ROOT)
    import bench
    groups = {"dft": ["dft_cols16", "dft_rows16", "dft_rows16_tma"], "linear_complexity": ["lc_kernel_win", "lc_kernel_reg", "lc_kernel_warp"],
              "rank": ["rank_kernel"], "pattern(serial,apen)": ["cyclic_hist16_kernel", "cyclic_hist_kernel"],
              "universal": ["universal_dist_kernel", "universal_sum_kernel", "universal_kernel"],
              "scan(freq,cusum,runs,excursions)": ["scan_kernel"],
              "blocks(blockfreq,longestrun,overlapping)": ["block_frequency_kernel", "longest_run_kernel", "overlapping_kernel"],
              "nonoverlapping": ["nonoverlapping_kernel"],
              "tails(p-values)": ["tails_slots_kernel", "tails_cusum_kernel", "tails_apen_kernel"]}
    pipe_of = {"linear_complexity": "alu", "rank": "alu", "pattern(serial,apen)": "lsu", "nonoverlapping": "lsu"}
    out = {"streams": B, "bits": n, "source": os.path.basename(raw), "source_sha256": bench.csrc_sha256(),
           "dram_bytes_per_group": {}, "pipe_util_pct": {}, "kernel_us_per_group": {}}
    for g, ks in groups.items():
        have = [k for k in ks if k in agg]
        if not have:
            continue
        us = sum(agg[k]["us"] for k in have)
        out["dram_bytes_per_group"][g] = sum(agg[k]["rd"] + agg[k]["wr"] for k in have)
        out["kernel_us_per_group"][g] = round(us, 1)
        if g in pipe_of and us > 0:
            out["pipe_util_pct"][g] = round(sum(agg[k][pipe_of[g]] for k in have) / us, 1)
    json.dump(out, open(sys.argv[4], "w"), indent=1)
print("\nsum of kernel durations: %.1f us for %d streams of %d bits -> %.1f Gbit/s if run back to back" % (
    tot, B, n, B * n / (tot * 1e-6) / 1e9))
