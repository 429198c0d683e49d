#!/usr/bin/env python
"""Kernel-level timing harness (development tool): per-stage CSA times, echo and GMTI throughput for a
list of sizes.  Prints one JSON line per measurement.  Usage: python tools/kbench.py [sizes...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import numpy as np
import torch

from nis_sar import device as dev, params, scenes

PEAK = 6549.1


def time_cuda(fn, iters=10, warm=3):
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


def csa_stages(n_az, n_rg, iters=10):
    prm = params.spaceborne_preset()
    plan = dev.CsaPlan(n_az, n_rg, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                       t_start=prm.t_start_fast)
    x = torch.view_as_complex(torch.randn((n_az, n_rg, 2), device="cuda"))
    out = torch.empty((n_rg, n_az), dtype=torch.complex64, device="cuda")
    total = time_cuda(lambda: plan.focus(x, out=out), iters)
    mx = torch.zeros(1, dtype=torch.float64, device="cuda")
    total_mx = time_cuda(lambda: plan.focus(x, out=out, max_sq=mx), iters)
    rec = {"what": "csa", "n_az": n_az, "n_rg": n_rg, "ms": total, "ms_with_max_sq": total_mx,
           "GBps_48B": 48.0 * n_az * n_rg / total * 1e-6, "frac": 48.0 * n_az * n_rg / total * 1e-6 / PEAK}
    try:
        plan.set_profiling(True)
        for _ in range(iters):
            plan.focus(x, out=out)
        st = {k: 0.0 for k in plan.STAGES}
        for b in range(iters):
            for k, v in plan.stage_times(b).items():
                st[k] += v / iters
        rec["stages_ms"] = st
        rec["stages_GBps"] = {k: 16.0 * n_az * n_rg / v * 1e-6 for k, v in st.items()}
    except Exception as e:  # general-size path has no stage events
        rec["stages_ms"] = str(e)[:60]
    plan.close()
    print(json.dumps(rec), flush=True)


def rda_vehicle(iters=5):
    """sar_vehicle_sim.py's own frame: 32768 pulses x 2048 samples, 1 us / 300 MHz chirp sampled at 360 MHz (361 taps)."""
    prm = params.airborne_vehicle_preset()
    n_pulses, n_ranges = 32768, 2048
    plan = dev.RdaPlan(n_pulses, n_ranges, lam=prm.Lambda, t_p=prm.T_p, kr=prm.k_rate, fs=360e6, prf=prm.PRF, vr=prm.V_sat,
                       range_grp=prm.R0)
    x = torch.view_as_complex(torch.randn((n_pulses, n_ranges, 2), device="cuda"))
    ms = time_cuda(lambda: plan.focus(x), iters)
    ms_all = time_cuda(lambda: plan.focus(x, want=plan.EXPORTS), iters)
    px = n_pulses * n_ranges
    print(json.dumps({"what": "rda_vehicle_frame", "n_pulses": n_pulses, "n_ranges": n_ranges, "ms_image_only": ms,
                      "Mpixel_per_s": px / ms * 1e-3, "frac_60B": 60.0 * px / ms * 1e-6 / PEAK, "ms_with_4_exports": ms_all}),
          flush=True)
    plan.close()


def rda(n_pulses, n_ranges, iters=10, t_p=10e-6):
    """Range-Doppler focusing.  Algorithmic bytes per pixel: range compression 16 + azimuth DFT 16 + RCMC/azimuth
    compression 16 + inverse DFT to magnitude 12 = 60 (image only; each exported map adds 8)."""
    prm = params.spaceborne_preset().replace(T_p=t_p)
    plan = dev.RdaPlan(n_pulses, n_ranges, lam=prm.Lambda, t_p=prm.T_p, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff,
                       range_grp=prm.R0)
    x = torch.view_as_complex(torch.randn((n_pulses, n_ranges, 2), device="cuda"))
    ms = time_cuda(lambda: plan.focus(x), iters)
    ms_all = time_cuda(lambda: plan.focus(x, want=plan.EXPORTS), iters)
    px = n_pulses * n_ranges
    print(json.dumps({"what": "rda", "n_pulses": n_pulses, "n_ranges": n_ranges, "taps": int(t_p * prm.FS) + 1,
                      "ms_image_only": ms, "Mpixel_per_s": px / ms * 1e-3, "GBps_60B": 60.0 * px / ms * 1e-6,
                      "frac": 60.0 * px / ms * 1e-6 / PEAK, "ms_with_4_exports": ms_all,
                      "GBps_92B": 92.0 * px / ms_all * 1e-6}), flush=True)
    plan.close()


def tdbp(n_pulses=2500, n_pix=512):
    """One VideoSAR frame of sar_batch_sim.py at its real size: 35-scatterer destroyer, 2500-pulse CPI, 22004-sample
    window, 12000-tap chirp, 512 x 512 pixels."""
    from nis_sar import api, targets as tg
    prm = params.batch_spotlight_preset()
    t_vec = (np.arange(n_pulses) - (n_pulses - 1) / 2) / prm.PRF
    pos_sat, vel_sat = scenes.orbit_trajectory(prm, t_vec, along="x")
    base = tg.generate_destroyer(center_pos=(0, 0, 0))
    l_ant = prm.Lambda * prm.R0 / 500.0
    raw, t0, n, vt = api.run_physics_spotlight(base, t_vec, pos_sat, vel_sat, 45.0, 15.0, l_ant, params=prm)
    ms_echo = time_cuda(lambda: api.run_physics_spotlight(base, t_vec, pos_sat, vel_sat, 45.0, 15.0, l_ant, params=prm), 3, 1)
    plan = dev.TdbpPlan(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, fs=prm.FS, t_start=t0, n_samples=n,
                        scene_size=500.0, nx=n_pix, ny=n_pix)
    rc = plan.range_compress(raw)
    ms_rc = time_cuda(lambda: plan.range_compress(raw), 5, 1)
    img = plan.backproject(rc, pos_sat, vel_sat, t_vec, vt)
    ms_bp = time_cuda(lambda: plan.backproject(rc, pos_sat, vel_sat, t_vec, vt, out=img), 5, 1)
    pairs = n_pulses * n_pix * n_pix
    print(json.dumps({"what": "tdbp_frame", "n_pulses": n_pulses, "n_samples": n, "pixels": n_pix * n_pix,
                      "ms_echo_incl_host_setup": ms_echo, "ms_range_compress": ms_rc, "ms_backproject": ms_bp,
                      "G_pixel_pulses_per_s": pairs / ms_bp * 1e-6,
                      "frames_per_s": 1e3 / (ms_echo + ms_rc + ms_bp)}), flush=True)
    plan.close()


def gmti(n):
    """K3 on a dense pair (noise: ~97 % of the pixels pass the 5 % threshold -- worst case for the compaction) and on a
    sparse one (one bright pixel); with and without max|slc1| handed in."""
    a = torch.view_as_complex(torch.randn((n, n, 2), device="cuda"))
    b = torch.view_as_complex(torch.randn((n, n, 2), device="cuda"))
    for tag in ("dense", "sparse"):
        if tag == "sparse":
            a[n // 2, n // 3] = 4000.0
        mx = torch.zeros(1, dtype=torch.float64, device="cuda")
        mx.fill_(float((torch.view_as_real(a).double() ** 2).sum(dim=-1).max()))
        ms = time_cuda(lambda: dev.gmti_fused(a, b, lazy=True, max_sq=mx), 20, 4)
        ms2 = time_cuda(lambda: dev.gmti_fused(a, b, lazy=True), 20, 4)
        ms3 = time_cuda(lambda: dev.gmti_fused(a, b, lazy=True, max_sq=mx, want=("ati_phase_masked",)), 20, 4)
        cnt = int(dev.gmti_fused(a, b, max_sq=mx, want=())["det_count"])
        print(json.dumps({"what": "gmti", "case": tag, "n": n, "detected_fraction": cnt / (n * n), "ms_all_products_max_given": ms,
                          "GBps_49B": 49.0 * n * n / ms * 1e-6, "frac": 49.0 * n * n / ms * 1e-6 / PEAK,
                          "ms_all_products_own_max": ms2, "ms_phase_and_detections_only": ms3,
                          "GBps_20B": 20.0 * n * n / ms3 * 1e-6}), flush=True)


def echo(kind):
    if kind == "stripmap8192":
        sc = scenes.stripmap_scene(8192, 8192)
        prm = sc["prm"]
        kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=prm.t_start_fast, fs=600e6, n_samples=8192)
        args = (sc["pos"], np.zeros(3), sc["rcs"], sc["pos_sat"], None, sc["t_vec"])
    elif kind.startswith("ati_default_"):
        sc = scenes.ati_scene(seed=0, num_pulses=int(kind[len("ati_default_"):-1]), num_clutter=5000, t_int=None)
        prm = sc["prm"]
        kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=prm.t_start_fast, fs=prm.FS, n_samples=13200)
        vhat = sc["vel_tx"] / np.linalg.norm(sc["vel_tx"], axis=1)[:, None]
        args = (sc["clutter_pos"], np.zeros(3), sc["clutter_rcs"], sc["pos_tx"], sc["pos_tx"] + vhat * sc["rx_offsets"][0], sc["t_vec"])
    else:
        sc = scenes.vehicle_scene(seed=0, num_pulses=1024, num_scatterers=20000)
        prm = sc["prm"]
        from oracle import sar_oracle as orc
        kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=orc.vehicle_window_start(prm.as_globals()),
                  fs=360e6, n_samples=2048)
        args = (sc["pos"], np.zeros(3), sc["rcs"], sc["pos_sat"], None, sc["t_vec"])
    out = dev.echo_accumulate(*args, **kw)
    ms = time_cuda(lambda: dev.echo_accumulate(*args, out=out, **kw), 3, 1)
    T, P, S = len(args[2]), len(args[5]), kw["n_samples"]
    print(json.dumps({"what": "echo", "kind": kind, "T": T, "P": P, "S": S, "ms": ms,
                      "G_updates_per_s_nominal": T * P * S / ms * 1e-6,
                      "Gsamples_per_s": P * S / ms * 1e-6}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["csa:4096x4096", "csa:8192x8192", "gmti:4096", "echo:stripmap8192", "echo:ati_default_256p",
                            "echo:vehicle"]
    for w in what:
        k, _, arg = w.partition(":")
        if k == "csa":
            a, b = arg.split("x")
            csa_stages(int(a), int(b))
        elif k == "tdbp":
            tdbp(*(int(v) for v in arg.split("x"))) if arg else tdbp()
        elif k == "rda_vehicle":
            rda_vehicle()
        elif k == "rda":       # rda:PULSESxSAMPLES[xPULSE_WIDTH_US]
            parts = arg.split("x")
            rda(int(parts[0]), int(parts[1]), t_p=float(parts[2]) * 1e-6 if len(parts) > 2 else 10e-6)
        elif k == "gmti":
            gmti(int(arg))
        elif k == "echo":
            echo(arg)
