#!/usr/bin/env python
"""Development tool: time the azimuth engines of the power-of-two CSA path (NIS_CSA_AZ knob: 0 = two-kernel four-step,
k = cluster configuration k) and check every engine against engine 0.  Usage: python tools/az_bench.py az | azone N K"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import torch

from nis_sar import device as dev, params

PEAK = 6549.1


def time_cuda(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run_az(n, ids):
    """cluster azimuth configurations (NIS_CSA_AZ) against the two-kernel four-step"""
    prm = params.spaceborne_preset()
    x = torch.view_as_complex(torch.randn((n, n, 2), device="cuda"))
    ref = None
    for az in ids:
        os.environ["NIS_CSA_AZ"] = str(az)
        try:
            plan = dev.CsaPlan(n, n, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                               t_start=prm.t_start_fast)
            out = torch.empty((n, n), dtype=torch.complex64, device="cuda")
            msq = torch.zeros(1, dtype=torch.float64, device="cuda")
            plan.focus(x, out=out, max_sq=msq)
            torch.cuda.synchronize()
        except Exception as e:
            print(json.dumps({"n": n, "az": az, "error": str(e)[:200]}), flush=True)
            continue
        if ref is None:
            ref = out.clone()
        err = float(torch.linalg.vector_norm(out - ref) / torch.linalg.vector_norm(ref))
        mx_ok = bool(abs(float(msq) - float((out.abs().double() ** 2).max())) <= 1e-6 * float(msq))
        ms = time_cuda(lambda: plan.focus(x, out=out))
        plan.set_profiling(True)
        for _ in range(5):
            plan.focus(x, out=out)
        st = plan.stage_times(0)
        print(json.dumps({"n": n, "az": az, "rel_l2_vs_az0": err, "max_ok": mx_ok, "ms": round(ms, 4),
                          "frac48": round(48.0 * n * n / ms * 1e-6 / PEAK, 4),
                          "stages": {k: round(v, 4) for k, v in st.items()}}), flush=True)
        plan.close()
    os.environ["NIS_CSA_AZ"] = "0"


if __name__ == "__main__":
    if sys.argv[1:2] == ["azone"]:
        run_az(int(sys.argv[2]), (int(sys.argv[3]),))
        sys.exit(0)
    if sys.argv[1:2] == ["az"]:
        for n, ids in ((1024, (0, 1)), (2048, (0, 1, 2)), (4096, (0, 1, 2, 3, 4)), (8192, (0, 1, 2, 3)), (16384, (0, 1))):
            run_az(n, ids)
        sys.exit(0)
    print("usage: az_bench.py az | azone N CONFIG")
