#!/usr/bin/env python
"""Segment the SASS of the first kernel in an ncu report into runs of similar execution count.

usage: python profiles/ncu_segments.py report.ncu-rep [units_per_launch]
Prints, per run: first SASS index, length, executions per unit (e.g. per scene), share of all warp
instructions.  Handy to see how often each phase of a kernel runs and what it costs.
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
    ins = []
    for r in rows[2:]:
        try:
            ins.append((r[ia], int(r[ie])))
        except (ValueError, IndexError):
            if r and r[0] == "Kernel Name" and ins:
                break
    tot = sum(n for _, n in ins) or 1
    print(f"{len(ins)} SASS instructions, {tot} warp instructions executed, {tot / units:.1f} per unit")
    i = 0
    while i < len(ins):
        j, s, base = i, 0, ins[i][1]
        while j < len(ins) and abs(ins[j][1] - base) <= 0.3 * max(base, 1):
            s += ins[j][1]
            j += 1
        if s > 0.004 * tot:
            print(f"idx {i:5d} len {j - i:4d} x {base / units:8.2f}/unit  {100 * s / tot:5.1f}%   {ins[i][0].strip()[:60]}")
        i = j


if __name__ == "__main__":
    main()
