#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.

usage: python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fname, seen = "", set()
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            if fname in seen:
                break            # second kernel instance
            seen.add(fname)
            continue
        if "Instructions Executed" in r:
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] != "":
            d = dict(zip(hdr, r))
            num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
            tag = {"common.cuh": "c", "raster_warp.cuh": "w"}.get(fname, fname[:8])
            lines.append((f"{tag}:{r[0]}", r[1], num(d["Instructions Executed"]), num(d["# Samples"]), d))
    tot_i = sum(l[2] for l in lines) or 1
    tot_s = sum(l[3] for l in lines) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    print("-- by instructions executed")
    for ln, src, n, s, _ in sorted(lines, key=lambda l: -l[2])[:top]:
        print(f"{ln:>7s} {100.0 * n / tot_i:5.1f}% inst {100.0 * s / tot_s:5.1f}% smp  {src.strip()[:110]}")
    print("-- by stall samples")
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for ln, src, n, s, d in sorted(lines, key=lambda l: -l[3])[:top]:
        st = sorted(((int(d[k]) if d[k] and d[k].isdigit() else 0, k[6:]) for k in keys), reverse=True)[:3]
        print(f"{ln:>7s} {100.0 * s / tot_s:5.1f}% smp {100.0 * n / tot_i:5.1f}% inst  "
              f"{' '.join(f'{k}:{v}' for v, k in st if v)}  | {src.strip()[:80]}")


if __name__ == "__main__":
    main()
