#!/usr/bin/env python3
"""Development aid: the handful of columns that matter from an `ncu --page raw --csv` dump, one block per launch.
usage: ncu_summary.py raw.csv [kernel-substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
pick = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "sm__maximum_warps_per_active_cycle_pct"]
for r in data:
    if len(sys.argv) > 2 and sys.argv[2] not in r[ki]:
        continue
    print("=====", r[ki][:70])
    for p in pick:
        if p in hdr:
            print("  %-72s %s %s" % (p, r[hdr.index(p)], units[hdr.index(p)]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
