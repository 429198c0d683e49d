#!/usr/bin/env python
"""Source-page digest of an .ncu-rep: the instructions of one kernel ranked by executed warp-instructions and by
pc-sampling stall samples, with their source lines (needs -lineinfo and --import-source on).
Usage: python tools/ncu_hot.py report.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

path = sys.argv[1]
kfilter = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
args = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"]
if kfilter:
    args += ["-k", "regex:" + kfilter]
out = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
per_line = defaultdict(lambda: [0, 0])
total_i = total_s = 0
insts = []
for r in rows:
    if hdr is None:
        if "Source" in r and any("Instructions Executed" in c for c in r):
            hdr = r
            i_src = r.index("Source")
            i_ex = next(i for i, c in enumerate(r) if c.strip() == "# Warp Instructions Executed" or c.strip() == "Instructions Executed")
            i_smp = next((i for i, c in enumerate(r) if c.strip().startswith("# Samples") or c.strip() == "Warp Stall Sampling (All Samples)"), None)
            i_loc = next((i for i, c in enumerate(r) if c.strip() in ("Location", "Source Location")), None)
        continue
    if len(r) <= i_ex:
        continue
    try:
        ex = int(float(r[i_ex].replace(",", "") or 0))
        sm = int(float(r[i_smp].replace(",", "") or 0)) if i_smp is not None else 0
    except ValueError:
        continue
    total_i += ex
    total_s += sm
    insts.append((ex, sm, r[i_src][:90], r[i_loc] if i_loc is not None else ""))
    if i_loc is not None:
        per_line[r[i_loc]][0] += ex
        per_line[r[i_loc]][1] += sm
print(f"# {path}: {len(insts)} SASS instructions, {total_i} warp-instructions executed, {total_s} stall samples")
if per_line:
    print("## by source line (share of executed warp-instructions, share of stall samples)")
    for loc, (ex, sm) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * ex / max(total_i, 1):6.2f}% {100.0 * sm / max(total_s, 1):6.2f}%  {loc}")
print("## hottest instructions by stall samples")
for ex, sm, src, loc in sorted(insts, key=lambda t: -t[1])[:top]:
    print(f"{100.0 * sm / max(total_s, 1):6.2f}% samples {100.0 * ex / max(total_i, 1):6.2f}% executed  {src}  [{loc}]")
