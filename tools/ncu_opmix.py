#!/usr/bin/env python
"""Opcode histogram (share of executed warp-instructions) of an .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
isrc, iex = h.index("Source"), h.index("Instructions Executed")
ops = {}
tot = 0
for r in rows[2:]:
    if len(r) > iex and r[iex].isdigit():
        parts = r[isrc].split()
        k = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
        k = k.split(".")[0] if "--full" not in sys.argv else k
        ops[k] = ops.get(k, 0) + int(r[iex])
        tot += int(r[iex])
print(f"total warp-instructions {tot}")
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:30]:
    print(f"  {k:24s} {100 * v / tot:6.2f}%")
