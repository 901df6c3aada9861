#!/usr/bin/env python
"""Summarises an .ncu-rep (captured on the GPU box with `ncu --set full --clock-control none --import-source on`)
into the handful of counters the roofline argument uses.  Usage: tools/ncu_summary.py REPORT.ncu-rep [--hot N]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {r[i]:>18s} {units[i]}")
    if "--hot" in sys.argv:
        n = int(sys.argv[sys.argv.index("--hot") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        h = rows[1]
        isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        data = [(r[isrc], int(r[isamp]), int(r[iex])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
        ts, te = sum(d[1] for d in data), sum(d[2] for d in data)
        print(f"  SASS instructions {len(data)}, warp-instructions executed {te}, stall samples {ts}")
        print("  hottest 64-instruction windows (share of stall samples / of executed instructions, top opcodes):")
        wins = []
        for b in range(0, len(data), 64):
            w = data[b:b + 64]
            ops = {}
            for d in w:
                parts = d[0].split()
                k = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
                ops[k] = ops.get(k, 0) + d[2]
            wins.append((sum(d[1] for d in w), sum(d[2] for d in w), b, sorted(ops.items(), key=lambda x: -x[1])[:5]))
        for s, e, b, top in sorted(wins, reverse=True)[:n]:
            print(f"    instr {b:5d}+64: samples {100 * s / max(ts, 1):5.1f}%  executed {100 * e / max(te, 1):5.1f}%  " +
                  " ".join(f"{k}:{100 * v / max(te, 1):.1f}%" for k, v in top))


if __name__ == "__main__":
    main()
