#!/usr/bin/env python
"""Regenerates profiles/r02z_sass_excerpt.txt from the in-tree library (cuobjdump -sass; run locally, no GPU)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pybatchrender_b200", "csrc", "libpbr_b200.so")
KERNEL = "_ZN3pbr18raster_warp_kernelILi14ELb1EEEvNS_8FrameDevE"


def sass(args):
    out = subprocess.run(["cuobjdump", "-sass"] + args + [LIB], capture_output=True, text=True).stdout
    return [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip() for l in out.splitlines() if not re.match(r"^\s*/\* 0x", l)]


whole = sass([])
ops = collections.Counter()
WANT = ("UBLKCP", "SYNCS", "BAR.SYNC", "ACQBULK", "UTMACMDFLUSH", "ATOMS", "VOTE", "REDUX", "PRMT", "BMSK", "FLO.U32", "ELECT",
        "HMMA", "IMMA", "UTCMMA", "TCGEN", "UTMALDG")
for l in whole:
    m = re.match(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and m.group(1).startswith(WANT):
        ops[".".join(m.group(1).split(".")[:4])] += 1
k = sass(["-fun", KERNEL])
out = ["# SASS evidence (cuobjdump -sass pybatchrender_b200/csrc/libpbr_b200.so, sm_100a), round-2 final build; made by profiles/sass_excerpt.py",
       "# 1-D bulk copies (TMA engine, no tensor map): UBLKCP.S.G = global->shared (background image, staged record chunks), UBLKCP.G.S = shared->global",
       "# (one per scene of a CTA); SYNCS.* = mbarrier arrive / try_wait; BAR.SYNC with a thread count = the named barrier of the shared geometry;",
       "# ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents; no tensor-core instruction anywhere (nothing on this path is",
       "# a contraction)", "", "## instruction counts over the whole library"]
out += [f"{n:7d} {op}" for op, n in ops.most_common()]
out += ["", "## raster_warp_kernel<14, true>: the bulk-copy / barrier / dependency-control instructions in program order"]
for i, l in enumerate(k):
    if re.search(r"UBLKCP|SYNCS|BAR\.SYNC|ACQBULK|UTMACMDFLUSH|PREEXIT", l):
        out.append(f"{i}: {l}")
# the sweep's record loop: from the FLO that finds the next record to the backward branch
start = next(i for i, l in enumerate(k) if "FLO.U32 R" in l and "IMAD R" in k[i + 1] and "-0x40" in k[i + 1])
end = next(i for i in range(start, len(k)) if re.search(r"@P\d BRA", k[i]))
out += ["", f"## raster_warp_kernel<14, true>: the sweep's loop over the records of a block, 32-bit depth keys ({end - start + 1} instructions per",
        "## (record, block) pair that covers a pixel; the branch behind VOTE.ANY skips the depth part when no lane is covered)"]
out += k[start:end + 1]
open(os.path.join(ROOT, "profiles", "r02z_sass_excerpt.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-(end - start + 8):][:12]))
