#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the design discussion uses."""
import csv
import re
import subprocess
import sys

PAT = re.compile(r'^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__registers_per_thread|launch__occupancy_limit_\w+|'
                 r'launch__grid_size|launch__block_size|sm__warps_active\.avg\.pct_of_peak_sustained_active|'
                 r'smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|'
                 r'sm__inst_executed_pipe_(fma|alu|lsu|xu|fmaheavy|fmalite|fp64|uniform)\.sum|'
                 r'sm__pipe_(fma|alu|xu|fmaheavy|fmalite)_cycles_active\.avg\.pct_of_peak_sustained_active|'
                 r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|'
                 r'smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__warps_eligible\.avg\.per_cycle_active|'
                 r'gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                 r'l1tex__throughput\.avg\.pct_of_peak_sustained_active|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                 r'smsp__pcsamp_warps_issue_stalled_\w+|sm__cycles_elapsed\.avg|l1tex__t_sectors_pipe_lsu_mem_local_op_(ld|st)\.sum)$')


def main(path, kernel_filter=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if kernel_filter and kernel_filter not in r[ki]:
            continue
        print("==", r[ki][:110], "block", r[hdr.index("Block Size")], "grid", r[hdr.index("Grid Size")])
        stalls = []
        for h, u, v in zip(hdr, units, r):
            if PAT.match(h):
                if "pcsamp_warps_issue_stalled" in h:
                    try:
                        stalls.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
                elif "issue_stalled" in h:
                    try:
                        fv = float(v.replace(",", ""))
                    except ValueError:
                        continue
                    if fv >= 0.05:
                        print(f"   {h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''):55s} {fv:8.2f}")
                else:
                    print(f"   {h:75s} {u:>10s} {v}")
        tot = sum(s for s, _ in stalls) or 1
        print("   pc-sampling stall mix:", ", ".join(f"{n}={100*s/tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
