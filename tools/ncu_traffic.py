#!/usr/bin/env python
"""Regenerates profiles/traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch.
Usage: tools/ncu_traffic.py KEY=REPORT.ncu-rep[:kernel-substring] ...   with KEY = workload/kind/channels/samples"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def traffic(rep, sub=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if sub is None or sub in r[ik]:
            return int(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]), r[ik]
    raise SystemExit(f"no kernel matching {sub!r} in {rep}")


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        rep, _, sub = rest.partition(":")
        b, name = traffic(rep, sub or None)
        table[key] = b
        print(f"{key}: {b} bytes  ({name[:60]})")
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
