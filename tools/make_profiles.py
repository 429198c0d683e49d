#!/usr/bin/env python
"""Turn the scratch artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <round-tag> [--launches F.csv] [--full F.ncu-rep ...] [--bench F.json] [--membench F.log ...]
"""
import argparse
import collections
import csv
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def launches_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | avg us | share of GPU time |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |")
    return "\n".join(out)


def dram_traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}[units[i]]
            tot += float(r[i].replace(",", "")) * scale
        res[name] = tot
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--full", nargs="*", default=[])
    ap.add_argument("--bench")
    ap.add_argument("--membench", nargs="*", default=[])
    ap.add_argument("--kbench", nargs="*", default=[])
    a = ap.parse_args()
    pdir = os.path.join(ROOT, "profiles")
    os.makedirs(pdir, exist_ok=True)
    if a.launches:
        with open(os.path.join(pdir, f"launches_{a.tag}.md"), "w") as f:
            f.write(f"# ncu launch list ({a.tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                    "`python bench.py --steps 2 --warmup 1` (cold-cache, serialised: compare shares, not absolutes).\n\n")
            f.write(launches_table(a.launches) + "\n")
        os.replace(a.launches, os.path.join(pdir, f"launches_{a.tag}.csv")) if False else None
        import shutil
        shutil.copy(a.launches, os.path.join(pdir, f"launches_{a.tag}.csv"))
    traffic = {}
    for rep in a.full:
        buf = io.StringIO()
        with redirect_stdout(buf):
            ncu_summary.main(rep)
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(pdir, f"{name}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on  ({rep})\n" + buf.getvalue())
        for k, v in dram_traffic(rep).items():
            if any(key in traffic for key in ("range",)) and "k_range" in k:
                continue          # first report wins (pass the bench-size capture first)
            if "k_range" in k:
                traffic["range"] = v
            elif "k_az_inner" in k:
                traffic["az_inner_fwd" if ", 0," in k or "false" in k else "az_inner_inv"] = v
            elif "k_az_outer_fwd" in k:
                traffic["az_outer_fwd"] = v
            elif "k_az_outer_inv" in k:
                traffic["az_outer_inv"] = v
            elif "k_echo" in k:
                traffic["echo"] = v
    if traffic:
        tp = os.path.join(pdir, "traffic.json")
        old = json.load(open(tp)) if os.path.isfile(tp) else {}
        old.update(traffic)
        json.dump(old, open(tp, "w"), indent=1)
    if a.bench:
        d = json.loads(open(a.bench).read().strip().splitlines()[-1])
        json.dump(d, open(os.path.join(pdir, f"bench_{a.tag}.json"), "w"), indent=1)
    for mb in a.membench + a.kbench:
        import shutil
        shutil.copy(mb, os.path.join(pdir, os.path.basename(mb).replace(".log", f"_{a.tag}.jsonl")))


if __name__ == "__main__":
    main()
