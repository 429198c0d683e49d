#!/usr/bin/env python
"""Development tool: print the interesting parts of a bench.py JSON line.  Usage: python tools/show_bench.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = lambda x, n=3: round(x, n) if isinstance(x, float) else x
print("value", r(d["value"], 0), d["unit"], "| ms/step", r(d["ms_per_step"]), "| e2e", r(d["e2e"]["value"], 0), r(d["e2e"]["ms_per_step"], 2), "ms | launches",
      d["gpu_launches"], "| n_gpus", d["n_gpus"])
ro = d["roofline"]
print("roofline", ro["kernel"], "frac", r(ro["frac"]), "launch_ms", r(ro["launch_ms"], 4), "| stages", {k: r(v, 4) for k, v in ro["stages_ms"].items()},
      "| csa_whole", r(ro["csa_whole"]["frac"]), r(ro["csa_whole"]["ms"], 4), "ms")
print("echo", {k: r(v) for k, v in d["echo"].items() if k != "bound"})
print("north_star", {k: r(v, 4) for k, v in (ro.get("north_star_frame") or {}).items() if k != "what"})
print("gmti", {k: r(v, 4) for k, v in (ro.get("gmti_stage") or {}).items() if k not in ("workload", "cpu_port")})
print("clocks", d.get("clocks"), "| numa", d["config"].get("numa"))
if "cpu_baseline" in d:
    print("cpu_baseline", r(d["cpu_baseline"]["value"], 4), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
for k, v in (d["config"].get("multi_gpu") or {}).items():
    print("--", k)
    for kk, vv in v.items():
        if kk in ("workload", "note", "what"):
            continue
        print("    ", kk, {a: r(b, 4) for a, b in vv.items() if a != "what"} if isinstance(vv, dict) else r(vv, 4))
for k, v in (d["config"].get("other_workloads") or {}).items():
    print("++", k, {a: (r(b, 4) if not isinstance(b, dict) else {x: r(y, 4) for x, y in b.items() if x not in ("what", "sample", "workload")})
                    for a, b in v.items() if a not in ("workload",)})
