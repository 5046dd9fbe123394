#!/usr/bin/env python
"""Summaries of ncu exports: launch list (gpu__time_duration per kernel) and raw-page key metrics.
  python tools/ncu_summary.py launches launches.csv
  ncu -i prof.ncu-rep --page raw --csv | python tools/ncu_summary.py raw -
"""
import collections
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        k = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-40s n=%3d total %9.1f us  share %5.1f%%  max %8.1f" % (k, len(v), sum(v), 100 * sum(v) / tot, max(v)))
    print("total us %.1f launches %d" % (tot, sum(len(v) for v in agg.values())))


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_sample_buffer_full"]


def raw(f):
    rows = list(csv.reader(f))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print("---")
        for w in WANT:
            if w in H:
                i = H.index(w)
                print("  %-66s %-14s %s" % (w, r[i][:60], U[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.stdin if sys.argv[2] == "-" else open(sys.argv[2]))
