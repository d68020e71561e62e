#!/usr/bin/env python
"""Builds profiles/<tag>_raster_warp_ncu.txt from the files one gpurun call leaves in gpurun_out/:
<tag>_bench.json, <tag>_bench_ref.json, <tag>_launches.csv, <tag>.ncu-rep (+ optional time-stamp logs).

usage: python profiles/make_summary.py r01n "title line" [timing_before.log timing_after.log]
"""
import collections, csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, title = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
out = [f"# {title}",
       "# command: ncu --set full --clock-control none --import-source on -k regex:raster_warp -s 6 -c 1 python bench.py --no-cpu-baseline --steps 16 --warmup 3",
       "",
       "## bench.py line of the same build (python bench.py: --steps 2000 --warmup 50)",
       open(f"{G}/{tag}_bench.json").read().strip(), "",
       "## reference arm on the same box (python bench.py --impl reference --steps 3 --warmup 1: the oracle port on all host cores)",
       open(f"{G}/{tag}_bench_ref.json").read().strip(), ""]
rows = list(csv.reader(open(f"{G}/{tag}_launches.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hi + 2:]:
    if len(r) > mv:
        try:
            agg[r[kn]].append(float(r[mv].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
out += ["## launch list (ncu --metrics gpu__time_duration.sum --clock-control none -c 300 over `bench.py --no-cpu-baseline --steps 64 --warmup 3`,",
        f"## raw file: {tag}_launches.csv; ns per launch, serialised and cold-cache, so absolute times are higher than in the bench).  The first 300",
        "## launches of the process: renderer construction (torch host math, one static-layer render) + warm-up and timed steps; this library's kernels:"]
own = {k: v for k, v in agg.items() if "pbr::" in k}
for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"  {len(v):4d} x {sum(v) / len(v):10.2f}   {100 * sum(v) / tot:5.1f}%  {k[:90]}")
ra = next((sum(v) / len(v) for k, v in own.items() if "raster_warp" in k), 0.0)
co = next((sum(v) / len(v) for k, v in own.items() if "compose" in k), 0.0)
if ra and co:
    out.append(f"  -> within a step: raster kernel {ra / 1000:.1f} us of {(ra + co) / 1000:.1f} us = {100 * ra / (ra + co):.0f} %, pose kernel "
               f"{100 * co / (ra + co):.0f} % (in the bench the raster prologue overlaps the pose kernel)")
out.append("")
raw = subprocess.run(["ncu", "-i", f"{G}/{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]
keys += [k for k in h if "issue_stalled" in k and "per_issue_active" in k]
out += ["## key metrics (one launch = 4096 scenes; dram__bytes_* is small because ncu profiles the launch in isolation and the",
        "## 50 MB of pixel writes are still dirty in the 126 MB L2 when the kernel ends)"]
for k in keys:
    if k in h:
        i = h.index(k)
        out.append(f"{k:95s} {v[i]:>16s} {u[i]}")
out.append("")
seg = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_segments.py"), f"{G}/{tag}.ncu-rep", "4096"], capture_output=True, text=True).stdout
out += ["## instruction mix by phase (profiles/ncu_segments.py, per scene)", seg.strip(), ""]
if len(sys.argv) > 4:
    out += ["## time stamps inside the kernel (PBR_W_TIMING build, %globaltimer per warp, us from the first CTA's entry)",
            "# before (the bulk stores of a CTA issued by its warp 0, which also owns a scene):", open(sys.argv[3]).read().strip(),
            "# after (the TMA engine driven by lane 0 of the first helper warp):", open(sys.argv[4]).read().strip()]
open(os.path.join(ROOT, "profiles", f"{tag}_raster_warp_ncu.txt"), "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_raster_warp_ncu.txt")
