#!/usr/bin/env python
"""Per-region stall breakdown of an .ncu-rep (source page, SASS): splits the kernel at its BAR.SYNC instructions (the
passes of a staged kernel) and prints, per region, executed warp-instructions, stall samples and the top stall reasons.
Usage: tools/ncu_stalls.py REPORT.ncu-rep [--min-share 1.0]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    regions, cur = [], {"start": 0, "ex": 0, "samp": 0, "st": {}, "ops": {}}
    n = 0
    for r in rows[2:]:
        if len(r) <= iex or not r[iex].isdigit():
            continue
        op = r[isrc].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        cur["ex"] += int(r[iex]); cur["samp"] += int(r[isamp])
        cur["ops"][op.split(".")[0]] = cur["ops"].get(op.split(".")[0], 0) + int(r[iex])
        for i, c in stall_cols:
            v = int(r[i]) if r[i].isdigit() else 0
            if v:
                cur["st"][c] = cur["st"].get(c, 0) + v
        n += 1
        if op.startswith("BAR") or op.startswith("EXIT"):
            cur["end"] = n
            regions.append(cur)
            cur = {"start": n, "ex": 0, "samp": 0, "st": {}, "ops": {}}
    cur["end"] = n
    regions.append(cur)
    tex, tsamp = sum(x["ex"] for x in regions), sum(x["samp"] for x in regions)
    min_share = float(sys.argv[sys.argv.index("--min-share") + 1]) if "--min-share" in sys.argv else 1.0
    print(f"total executed {tex}, samples {tsamp}")
    for x in regions:
        if 100.0 * x["samp"] / max(tsamp, 1) < min_share:
            continue
        top = sorted(x["st"].items(), key=lambda kv: -kv[1])[:5]
        ops = sorted(x["ops"].items(), key=lambda kv: -kv[1])[:4]
        print(f"  sass {x['start']:5d}-{x['end']:5d}: exec {100.0 * x['ex'] / tex:5.1f}%  samples {100.0 * x['samp'] / tsamp:5.1f}%  "
              f"ratio {x['samp'] / tsamp / max(x['ex'] / tex, 1e-9):4.2f}  " + " ".join(f"{k[6:]}:{100.0 * v / max(x['samp'], 1):.0f}%" for k, v in top)
              + "  | " + " ".join(f"{k}:{100.0 * v / max(x['ex'], 1):.0f}%" for k, v in ops))


if __name__ == "__main__":
    main()
