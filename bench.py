#!/usr/bin/env python
"""bench.py -- throughput of the SAR hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1]): spaceborne stripmap point-target grid, 8192 x 8192 raw echo
synthesis (sar_satellite_sim.py echo model, 9 x 9 scatterer grid) followed by Chirp Scaling focusing
of the 8192 x 8192 frame.  One step = echo synthesis + CSA focusing of one frame.  With --gpus N
every rank focuses its own frame (VideoSAR-style frame parallelism, no data-path collective):
weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

prints ONE JSON line.  `value` is Mpixel/s of focused image over all ranks with inputs resident in
HBM; `e2e` is the same step through the drop-in Python API with host buffers (scene arrays copied
host->device, the complex128 focused image copied device->host, every step).

At every N the same run also measures the three multi-GPU workloads BASELINE.json names -- nested under
`config.multi_gpu` so that they survive in the driver's record:
  config 4  VideoSAR: 64 two-channel 4096 x 4096 frames (CSA x2 + fused DPCA/ATI), round robin over ranks  -> frames/s
            (strong scaling) and the per-GPU fraction of the HBM roofline (the north_star target);
  config 3  dense vehicle scene: 1e5 scatterers x 32768 pulses x 2048 samples sharded by scatterer, partial echoes
            reduced by NCCL (torch.distributed and the library's own nis_echo_reduce) or inside the synthesis kernel
            (system-scope RED.ADD into the owner's HBM over NVLink);
  config 5  HRWS: one 4096 x 4096 receive channel per rank, pair (k, k+1) formed after an NCCL ring shift
            (isend/irecv, nis_slc_exchange) or by the DPCA/ATI kernel reading the neighbour's image over NVLink;
each with an in-run equality check against a one-rank recomputation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "nis-sar-amtigmti-video_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_AZ = N_RG = 8192
GRID_SIDE = 9
METRIC = "csa_focused_mpixels_per_s"
UNIT = "Mpixel/s"
CSA_ALGO_BYTES_PER_PIXEL = 48.0      # 3 fused passes x (8 B read + 8 B write), SURVEY.md section 8d
STAGE_ALGO_BYTES_PER_PIXEL = 16.0    # one kernel: read + write one complex64


def workload(n_az=N_AZ, n_rg=N_RG):
    from nis_sar import scenes
    sc = scenes.stripmap_scene(num_pulses=n_az, num_samples=n_rg, n_side=GRID_SIDE, half_extent=1000.0)
    return sc


def config_dict(args, extra=None):
    cfg = {"workload": f"stripmap point-target grid {GRID_SIDE}x{GRID_SIDE}, {N_AZ}x{N_RG} raw echo (monostatic, "
                       f"sar_satellite_sim.py model) + CSA focus; one frame per rank per step",
           "n_az": N_AZ, "n_rg": N_RG, "scatterers": GRID_SIDE * GRID_SIDE,
           "l2_policy": "inputs larger than L2 (537 MB workspace per frame vs 126 MB L2)",
           "parallelism": f"frame-parallel x{args.gpus}"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------ CPU arms
# The reference's CPU path for this step is two single-threaded numpy functions (run_physics_engine,
# sar_focus_csa).  A full 8192 x 8192 step costs ~450 s of echo synthesis + ~40 s (21 GB) of CSA per core, so the
# CPU arms time a BOUNDED SAMPLE and extrapolate linearly, and say so: `kind` = "port, extrapolated".
#   echo: `pulses` whole pulses of the 8192-pulse aperture (cost is exactly linear in pulses);
#   CSA : one full 4096 x 4096 frame (a quarter of the pixels of the bench frame; the reference's cost per pixel grows
#         ~ log N and with cache misses, so this sample FLATTERS the CPU by an estimated 10-20 %).
REF_ECHO_PULSES = 64
REF_CSA_N = 4096
REF_CSA_GB_PER_PROC = 7.0          # peak RSS of the numpy CSA at 4096^2 (complex128 temporaries), with margin


def _cpu_echo_rate(sc, pulses):
    """Oracle (numpy port of run_physics_engine) on a pulse subset: scatterer-samples per second."""
    from oracle import sar_oracle as orc
    prm = sc["prm"]
    idx = np.linspace(0, len(sc["t_vec"]) - 1, pulses).astype(int)
    t0 = time.perf_counter()
    orc.echo_monostatic(sc["pos"], sc["rcs"], sc["t_vec"][idx], sc["pos_sat"][idx], prm.as_globals(), n_samples=N_RG)
    dt = time.perf_counter() - t0
    return len(sc["rcs"]) * pulses * N_RG / dt, dt


def _cpu_csa_rate(n, seed=0):
    from oracle import sar_oracle as orc
    from nis_sar import params
    prm = params.spaceborne_preset()
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    t0 = time.perf_counter()
    orc.focus_csa(x, prm.Lambda, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, prm.t_start_fast)
    dt = time.perf_counter() - t0
    return n * n / dt, dt


def _cpu_worker(args):
    seed, pulses, n_csa = args
    sc = workload()
    e_rate, e_dt = _cpu_echo_rate(sc, pulses)
    c_rate, c_dt = _cpu_csa_rate(n_csa, seed)
    return e_rate, c_rate, e_dt + c_dt


def _cpu_procs():
    """Host processes the CPU arm may use: every core it is allowed to run on, capped by memory (7 GB per numpy CSA)."""
    cores = max(1, len(os.sched_getaffinity(0)))
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
    except (ValueError, OSError):
        avail_gb = 64.0
    return max(1, min(cores, int(avail_gb * 0.8 / REF_CSA_GB_PER_PROC))), cores


def cpu_step_throughput(procs, pulses=REF_ECHO_PULSES, n_csa=REF_CSA_N):
    """Modelled Mpixel/s of the full step (echo + CSA) on `procs` host processes, each running the single-threaded numpy
    port on its own bounded sample.  Per-pixel CPU cost = scatterers / echo_rate + 1 / csa_rate; a process's modelled
    step time = 8192 x 8192 x that cost."""
    import multiprocessing as mp
    if procs <= 1:
        res = [_cpu_worker((0, pulses, n_csa))]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, [(i, pulses, n_csa) for i in range(procs)])
    T = GRID_SIDE * GRID_SIDE
    per_proc = [1.0 / (T / e + 1.0 / c) for e, c, _ in res]
    return sum(per_proc) / 1e6, float(np.mean([e for e, _, _ in res])), float(np.mean([c for _, c, _ in res])), \
        float(max(d for _, _, d in res))


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the step (numpy port, the reference itself being Python
    that cannot be imported -- DESIGN.md section 1) on all the host cores memory allows.  Every step times a bounded
    sample (64 of 8192 pulses + one 4096^2 CSA per process) and converts it into the MODELLED time of one full step;
    the line says so (`kind`, `sample`, `modelled_ms_per_step`) and also carries the wall time actually spent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs, cores = _cpu_procs()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    budget_s = float(os.environ.get("NIS_REF_BUDGET_S", "170"))
    vals, sample, t_all0 = [], None, time.perf_counter()
    n_total = args.warmup + args.steps
    for i in range(n_total):
        t0 = time.perf_counter()
        v, e_rate, c_rate, dt = cpu_step_throughput(procs)
        wall = time.perf_counter() - t0
        if i >= args.warmup or i == n_total - 1:
            vals.append((v, wall))
        sample = (f"{procs} processes ({cores} cores visible; 7 GB per process), each: numpy port of run_physics_engine on "
                  f"{REF_ECHO_PULSES} of {N_AZ} pulses x {N_RG} samples x {GRID_SIDE * GRID_SIDE} scatterers ({e_rate:.3g} "
                  f"scatterer-samples/s) + sar_focus_csa port on ONE {REF_CSA_N}x{REF_CSA_N} frame ({c_rate / 1e6:.3g} Mpixel/s); "
                  f"per-pixel costs added and scaled to the {N_AZ}x{N_RG} step (extrapolation, not a run of the full step)")
        # keep the whole arm inside a few minutes: stop early when the next sample would not fit (steps_run says how many ran)
        if time.perf_counter() - t_all0 + wall > budget_s and len(vals) >= 1:
            break
    v = float(np.mean([x for x, _ in vals]))
    pixels = float(N_AZ) * N_RG
    modelled_ms = 1e3 * pixels / (v / procs * 1e6)          # one process, one full frame
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_run": len(vals), "warmup": args.warmup,
            "ms_per_step": modelled_ms / procs,              # modelled: `procs` frames in flight, one finishing every ...
            "modelled_ms_per_step_one_core": modelled_ms,
            "sample_wall_ms": 1e3 * float(np.mean([d for _, d in vals])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port, extrapolated", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- multi-GPU workloads
def _ev_pair():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_ms(ms, device, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _all_true(flag, device, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def synthetic_collection(device, rows, n, seed, n_side=3):
    """[rows, n] complex64 raw data: unit-variance receiver noise (seeded; same generator on the same GPU model, so every
    rank can regenerate any channel) + the echoes of an n_side x n_side grid of unit-RCS point scatterers synthesised by K1 on
    the stripmap geometry of configs[1].  After focusing the scatterers stand ~70 dB above the noise floor, so the 5 %
    detection mask is selective, as in a real scene (pure noise would flag ~97 % of the pixels)."""
    import torch
    from nis_sar import device as dev, scenes
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.view_as_complex(torch.randn((rows, n, 2), generator=gen, device=device))
    sc = scenes.stripmap_scene(num_pulses=rows, num_samples=n, n_side=n_side, half_extent=300.0)
    prm = sc["prm"]
    # an n-sample window is shorter than the 20 us chirp: open it in the middle of the scene-centre echo, so that every sample
    # carries signal (the reference's own window formula, :112, would start it before the echo arrives)
    t_start = 2 * prm.R0 / prm.C + prm.T_p / 2 - (n / 600e6) / 2
    dev.echo_accumulate(sc["pos"], np.zeros(3), sc["rcs"], sc["pos_sat"], None, sc["t_vec"], c=prm.C, fc=prm.FC, k_rate=prm.k_rate,
                        t_p=prm.T_p, t_start=t_start, fs=600e6, n_samples=n, device=device, out=x, accumulate=True)
    return x, prm, t_start


def bench_config4_videosar(device, rank, world, peak_gbs, n_frames=64, n=4096, stride=8, keep_pair=False):
    """BASELINE.json configs[3]: a VideoSAR sub-aperture sequence -- `n_frames` two-channel n x n frames cut from one long
    seeded collection with the stride pattern of sar_batch_sim.py:303-310, frame f = pulses [stride f, stride f + n] of each
    receive channel -- focused frame-parallel (round robin over ranks, no data-path collective): per frame CSA of rx1[1:]
    and rx2[:-1] (the DPCA pulse shift, sar_ati_dcpa_sim_csa.py:402-403) on two streams + the fused DPCA/ATI/detection
    kernel with every product.  STRONG scaling: the sequence is fixed, frames/s = n_frames / max-over-ranks time.
    Eager launches (no CUDA graph).  The 16-byte detection record of every frame is gathered and rank 0 recomputes a
    frame another rank owned."""
    import torch
    import torch.distributed as dist
    from nis_sar import device as dev, dist as nd, params
    rows = n + 1 + stride * (n_frames - 1)
    coll = []
    for c in range(2):
        x, prm, t_start = synthetic_collection(device, rows, n, 404 + c)
        coll.append(x)
    mk = lambda: dev.CsaPlan(n, n, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                             t_start=t_start, device=device)
    pa, pb = mk(), mk()
    s1 = torch.empty((n, n), dtype=torch.complex64, device=device)
    s2 = torch.empty_like(s1)
    mx = torch.zeros(1, dtype=torch.float64, device=device)
    st_a, st_b = torch.cuda.Stream(device), torch.cuda.Stream(device)
    gbuf = dev.GmtiBuffers((n, n), device=device)                # products, detection list, workspace: allocated once
    records = torch.zeros((n_frames, 4), dtype=torch.int32, device=device)
    rec_view = gbuf.result.view(torch.int32)

    def frame(f, two_streams):
        cur = torch.cuda.current_stream(device)
        if two_streams:      # the channels are independent until the pairing: their HBM-bound and compute-bound kernels overlap
            st_a.wait_stream(cur)
            st_b.wait_stream(cur)
            with torch.cuda.stream(st_a):
                pa.focus(coll[0][stride * f + 1: stride * f + 1 + n], out=s1, max_sq=mx)
            with torch.cuda.stream(st_b):
                pb.focus(coll[1][stride * f: stride * f + n], out=s2)
            cur.wait_stream(st_a)
            cur.wait_stream(st_b)
        else:
            pa.focus(coll[0][stride * f + 1: stride * f + 1 + n], out=s1, max_sq=mx)
            pb.focus(coll[1][stride * f: stride * f + n], out=s2)
        out = dev.gmti_fused(s1, s2, max_sq=mx, lazy=True, buffers=gbuf)
        records[f].copy_(rec_view, non_blocking=True)
        return out
    mine = list(nd.frame_indices(n_frames))
    cur = torch.cuda.current_stream(device)
    timing = {}
    for tag, two in (("eager_two_streams", True), ("eager_one_stream", False)):
        for f in mine[:3]:
            frame(f, two)
        records.zero_()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        e0, e1 = _ev_pair()
        e0.record(cur)
        for f in mine:
            frame(f, two)
        e1.record(cur)
        torch.cuda.synchronize(device)
        timing[tag] = (e0.elapsed_time(e1), _max_ms(e0.elapsed_time(e1), device, world))
    best = min(timing, key=lambda k: timing[k][1])
    my_ms, ms = timing[best]
    if world > 1:
        dist.all_reduce(records)                                 # every frame was written by exactly one rank
    got = records.cpu().numpy().copy()
    equal = None
    if rank == 0:
        f_chk = 1 if world > 1 else n_frames - 1                 # owned by rank 1 when there is one
        frame(f_chk, False)
        torch.cuda.synchronize(device)
        equal = bool((records[f_chk].cpu().numpy() == got[f_chk]).all())
    # the same frame replayed from a CUDA graph (fixed input pointers: what a production loop with a staging buffer does)
    graph_ms = None
    try:
        frame(mine[0], True)
        torch.cuda.synchronize(device)
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            frame(mine[0], True)
        for _ in range(3):
            gph.replay()
        torch.cuda.synchronize(device)
        e0, e1 = _ev_pair()
        e0.record(cur)
        for _ in range(30):
            gph.replay()
        e1.record(cur)
        torch.cuda.synchronize(device)
        graph_ms = e0.elapsed_time(e1) / 30
        del gph
    except Exception:
        graph_ms = None
    fr_bytes = (2 * CSA_ALGO_BYTES_PER_PIXEL + 49.0) * n * n
    per_gpu_ms = my_ms / max(len(mine), 1)
    pa.close()
    pb.close()
    del coll
    pair = (s1, s2) if keep_pair else None
    if not keep_pair:
        del s1, s2
    torch.cuda.empty_cache()
    return {"focused_pair": pair,
            "workload": f"{n_frames} two-channel {n}x{n} frames (sub-apertures at stride {stride} of one seeded collection): CSA x2 "
                        f"+ fused DPCA/ATI/threshold/compaction, all products; round robin over {world} rank(s); eager launches (headline) "
                        f"and one frame replayed from a CUDA graph",
            "frames": n_frames, "ms_total": ms, "frames_per_s": n_frames / (ms * 1e-3), "scaling": "strong",
            "launch_mode": best, "ms_total_by_launch_mode": {k: v[1] for k, v in timing.items()},
            "ms_per_frame_cuda_graph_replay": graph_ms,
            "frac_of_hbm_peak_cuda_graph_replay": (fr_bytes / (graph_ms * 1e-3) / 1e9 / peak_gbs) if graph_ms else None,
            "ms_per_frame_per_gpu": per_gpu_ms, "algorithmic_bytes_per_frame": fr_bytes,
            "per_gpu_achieved_GBps": fr_bytes / (per_gpu_ms * 1e-3) / 1e9,
            "per_gpu_frac_of_hbm_peak": fr_bytes / (per_gpu_ms * 1e-3) / 1e9 / peak_gbs,
            "collective": "none in the data path (16-byte detection records all-reduced after the timed region)",
            "detections_frame0": int(got[0][0]), "detected_fraction_frame0": float(got[0][0]) / (n * n),
            "equals_one_rank_recompute": equal,
            "note": "4096^2 x 8 B = 134 MB per array vs 126 MB of L2: intermediate passes partly hit L2 (ncu: 227 MB of DRAM "
                    "traffic for a 268 MB algorithmic pass), so the fraction is against HBM peak, not a pure-DRAM figure"}


def bench_config3_scatterer_shards(device, rank, world, num_scatterers=100000, num_pulses=32768):
    """BASELINE.json configs[2]: sar_vehicle_sim.py geometry (fs 360 MHz, 2048 samples, 32768 pulses: :43, :85-89) with
    1e5 scatterers tiled from vehicle_targets.py, sharded BY SCATTERER over the ranks; the partial [P, S] echoes (512 MB)
    meet in one of three ways: torch.distributed all-reduce (NCCL), the library's nis_echo_reduce (NCCL, C ABI), or the
    synthesis kernel's own epilogue adding each pulse block into its owner's HBM with system-scope RED.ADD over NVLink
    (reduce-scatter semantics, no partial echo materialised).  Checked against a one-rank recomputation of 16 pulses."""
    import torch
    import torch.distributed as dist
    from nis_sar import device as dev, dist as nd, scenes
    sc = scenes.vehicle_scene(seed=0, num_pulses=num_pulses, num_scatterers=num_scatterers)
    prm = sc["prm"]
    S = 2048
    kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=(2 * prm.R0 / prm.C) - (S / 360e6) / 2,
              fs=360e6, n_samples=S, device=device)
    zero3 = np.zeros(3)
    part = torch.zeros((num_pulses, S), dtype=torch.complex64, device=device)

    def shard(a, b):
        return dev.echo_accumulate(sc["pos"][a:b], zero3, sc["rcs"][a:b], sc["pos_sat"], None, sc["t_vec"], out=part, **kw)

    def shard_into(a, b, p0, p1, dst):
        dev.echo_accumulate(sc["pos"][a:b], zero3, sc["rcs"][a:b], sc["pos_sat"], None, sc["t_vec"], out=dst,
                            pulse_range=(p0, p1), accumulate="atomic", **kw)
    res = {"workload": f"{num_scatterers} scatterers x {num_pulses} pulses x {S} samples (sar_vehicle_sim.py geometry), sharded by "
                       f"scatterer over {world} rank(s); partial echoes 512 MB per rank",
           "scatterer_samples": float(num_scatterers) * num_pulses * S}
    comm = nd.CComm(device=device) if world > 1 else None
    t0, t1 = nd.block_range(num_scatterers, rank, world)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    cur = torch.cuda.current_stream(device)
    routes = [("nccl_torch", None)] + ([("nccl_c_abi", comm)] if comm is not None else [])
    ref_block = None
    for name, cm in routes:
        for rep in range(2):                                      # first pass warms the communicator
            torch.cuda.synchronize(device)
            if world > 1:
                dist.barrier()
            e0.record(cur)
            shard(t0, t1)
            e1.record(cur)
            if world > 1:
                if cm is not None:
                    cm.echo_reduce(part)
                else:
                    dist.all_reduce(torch.view_as_real(part))
            e2.record(cur)
            torch.cuda.synchronize(device)
        tot, red = _max_ms(e0.elapsed_time(e2), device, world), _max_ms(e1.elapsed_time(e2), device, world)
        res[name] = {"ms_total": tot, "ms_synthesis": tot - red, "ms_reduce": red,
                     "g_scatterer_samples_per_s": res["scatterer_samples"] / (tot * 1e-3) / 1e9}
        if ref_block is None:
            p0, p1 = nd.block_range(num_pulses, rank, world)
            ref_block = part[p0:p1].clone()
    res["nccl_routes_bit_equal"] = None
    if comm is not None:
        res["nccl_routes_bit_equal"] = _all_true(torch.equal(part[p0:p1], ref_block), device, world)
    # one-rank recomputation of 16 pulses of this rank's own block, all scatterers
    chk = torch.zeros((num_pulses, S), dtype=torch.complex64, device=device) if world > 1 else None
    p0, p1 = nd.block_range(num_pulses, rank, world)
    rows = slice(p0, min(p0 + 16, p1))
    if world > 1:
        dev.echo_accumulate(sc["pos"], zero3, sc["rcs"], sc["pos_sat"], None, sc["t_vec"], out=chk,
                            pulse_range=(rows.start, rows.stop), **kw)
        err = float(torch.linalg.vector_norm(ref_block[: rows.stop - rows.start] - chk[rows]) /
                    torch.linalg.vector_norm(chk[rows]))
        res["rel_l2_vs_one_rank_recompute"] = _max_ms(err, device, world)
        res["equals_one_rank_recompute"] = res["rel_l2_vs_one_rank_recompute"] < 1e-5     # fp32 summation order differs
        del chk
    else:
        res["equals_one_rank_recompute"] = True
    # fused route: RED.ADD into the owner's rows over NVLink
    try:
        shared = nd.SharedBuffer((num_pulses, S), torch.complex64, device=device)
        res["peer_native_atomics"] = bool(shared.native_atomics)
        for rep in range(2):
            torch.cuda.synchronize(device)
            if world > 1:
                dist.barrier()
            e0.record(cur)
            own = nd.echo_scatterer_shards_p2p(shard_into, num_scatterers, shared)
            e1.record(cur)
            torch.cuda.synchronize(device)
        tot = _max_ms(e0.elapsed_time(e1), device, world)
        blk = shared.local[own[0]:own[1]]
        err = float(torch.linalg.vector_norm(blk - ref_block) / torch.linalg.vector_norm(ref_block))
        res["fused_red_add"] = {"ms_total": tot, "g_scatterer_samples_per_s": res["scatterer_samples"] / (tot * 1e-3) / 1e9,
                                "rel_l2_vs_nccl_route": _max_ms(err, device, world),
                                "what": "every rank synthesises its scatterers for all pulse blocks; the kernel epilogue adds "
                                        "block b into rank b's HBM (RED.E.ADD.F32.SYS over NVLink); barriers included"}
        shared.close()
    except Exception as e:                                       # no peer path on this box: the NCCL routes stand
        res["fused_red_add"] = {"error": str(e)[:200]}
    if comm is not None:
        comm.close()
    best = min((v["ms_total"], k) for k, v in res.items() if isinstance(v, dict) and "ms_total" in v)
    res["fastest_route"] = best[1]
    r = res.get("nccl_torch", {})
    res["limiter"] = ("synthesis (FP32 pipe): the reduce is %.1f %% of the step" % (100 * r["ms_reduce"] / r["ms_total"])) \
        if world > 1 else "synthesis (FP32 pipe); no exchange at 1 rank"
    del part, ref_block
    torch.cuda.empty_cache()
    return res


def bench_config5_hrws(device, rank, world, n=4096):
    """BASELINE.json configs[4]: HRWS -- one n x n receive channel per rank (seeded, so any rank can regenerate any
    channel), focused where it lives; DPCA/ATI pair (k, k+1) is formed on rank k from its own image and the neighbour's,
    which arrives by an NCCL ring shift (torch isend/irecv; the library's nis_slc_exchange) or is read in place over NVLink
    by the products kernel (peer-mapped buffer).  At one rank both channels are local (no exchange)."""
    import torch
    import torch.distributed as dist
    from nis_sar import device as dev, dist as nd, params
    prm = params.spaceborne_preset().replace(n_samples=n, window_s=n / 600e6)
    plan = dev.CsaPlan(n, n, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                       t_start=2 * prm.R0 / prm.C + prm.T_p / 2 - (n / 600e6) / 2, device=device)

    def channel(k):
        x, _, _ = synthetic_collection(device, n, n, 9000 + k)
        return plan.focus(x).clone()
    mine = channel(rank)
    prod = lambda a, b: dev.gmti_fused(a, b, lazy=True)          # every product, no host read-back
    res = {"workload": f"{max(world, 2)} receive channels of {n}x{n}, one per rank; {max(world - 1, 1)} adjacent DPCA/ATI pair(s), "
                       f"all products + detection list", "pairs": max(world - 1, 1)}
    e0, e1 = _ev_pair()
    cur = torch.cuda.current_stream(device)

    def timed(fn, reps=5):
        for _ in range(3):           # the caching allocator needs two rounds to recycle the ~750 MB of product buffers
            out = fn()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        evs = [_ev_pair() for _ in range(reps)]
        for a_ev, b_ev in evs:       # per-repetition events, median: one slow NCCL send / recv (lazy connection set-up on a
            a_ev.record(cur)         # ring neighbour) otherwise decides the mean of five
            out = fn()
            b_ev.record(cur)
        torch.cuda.synchronize(device)
        return _max_ms(float(np.median([a_ev.elapsed_time(b_ev) for a_ev, b_ev in evs])), device, world), out
    if world == 1:
        other = channel(1)
        ms, out = timed(lambda: prod(mine, other))
        res["local_pair"] = {"ms": ms}
        res["equals_one_rank_recompute"] = True
        res["limiter"] = "none: both channels on one GPU"
        plan.close()
        return res
    comm = nd.CComm(device=device)
    keep = {}

    def via_c():
        nxt = comm.slc_exchange(mine)
        return None if nxt is None else prod(mine, nxt)
    for name, fn in (("nccl_torch_isend_irecv", lambda: nd.pair_products(mine, prod)), ("nccl_c_abi_slc_exchange", via_c)):
        ms, out = timed(fn)
        res[name] = {"ms_per_pair_step": ms}
        if out is not None:
            keep[name] = (out["result_dev"].clone(), out["det_idx_raw"][:4096].clone())
    try:
        sh = nd.SharedBuffer((n, n), torch.complex64, device=device)
        sh.local.copy_(mine)
        ms, out = timed(lambda: nd.pair_products_p2p(sh, prod))
        res["peer_read_over_nvlink"] = {"ms_per_pair_step": ms,
                                        "what": "k_gmti_fused loads slc2 (8 of its 16 input bytes per pixel) from the neighbour's HBM"}
        if out is not None:
            keep["peer_read_over_nvlink"] = (out["result_dev"].clone(), out["det_idx_raw"][:4096].clone())
        sh.close()
    except Exception as e:
        res["peer_read_over_nvlink"] = {"error": str(e)[:200]}
    ok = True
    if rank < world - 1:
        ref = prod(mine, channel(rank + 1))                      # one-rank recomputation of this rank's pair
        torch.cuda.synchronize(device)
        n_det = min(int(ref["result_dev"].view(torch.int32)[0].item()), 4096)    # entries past det_count are not written
        for name, (rec, idx) in keep.items():
            ok = ok and bool(torch.equal(rec, ref["result_dev"])) and bool(torch.equal(idx[:n_det], ref["det_idx_raw"][:n_det]))
        res["detections_pair0"] = int(ref["result_dev"].view(torch.int32)[0].item()) if rank == 0 else None
    res["equals_one_rank_recompute"] = _all_true(ok, device, world)
    a = res.get("peer_read_over_nvlink", {}).get("ms_per_pair_step")
    b = res["nccl_torch_isend_irecv"]["ms_per_pair_step"]
    res["limiter"] = (f"the ring shift of one 134 MB image (NCCL send/recv: {b:.2f} ms per step) -- removed by the in-place peer read "
                      f"({a:.2f} ms)") if a else f"the ring shift of one 134 MB image ({b:.2f} ms per step)"
    comm.close()
    plan.close()
    del mine
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nis_sar import api, device as dev, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_info = {"policy": "none"}
    if world > 1:
        # one process per GPU on a multi-socket host: keep each rank (and the pinned result buffers it allocates) on one
        # NUMA node -- the GPU's own, or spread over the nodes when the platform reports every GPU on the same one
        from nis_sar import hostio
        numa_info = hostio.bind_rank_to_numa(local, world, policy=os.environ.get("NIS_NUMA_POLICY", "auto"))
        dist.init_process_group("nccl", device_id=device)

    sc = workload()
    prm = sc["prm"]
    targets = [{"position": p, "rcs": r} for p, r in zip(sc["pos"], sc["rcs"])]
    echo_kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=prm.t_start_fast, fs=600e6,
                   n_samples=N_RG, device=device)
    plan = dev.cached_plan(N_AZ, N_RG, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff,
                           r_ref=prm.R0, t_start=prm.t_start_fast, device=device)
    plan.set_profiling(True)
    raw = torch.zeros((N_AZ, N_RG), dtype=torch.complex64, device=device)
    slc = torch.empty((N_RG, N_AZ), dtype=torch.complex64, device=device)
    # scene arrays resident in HBM for the device-timed region
    import ctypes as C
    pos0_d = torch.from_numpy(np.ascontiguousarray(sc["pos"])).to(device)
    vel_d = torch.zeros(3, dtype=torch.float64, device=device)
    amp_d = torch.from_numpy(np.sqrt(sc["rcs"])).to(device)
    ptx_d = torch.from_numpy(np.ascontiguousarray(sc["pos_sat"])).to(device)
    ts_d = torch.from_numpy(np.ascontiguousarray(sc["t_vec"])).to(device)
    tf_d = torch.from_numpy(dev.fast_time_axis(prm.t_start_fast, N_RG, 600e6)).to(device)
    eprm = _lib.EchoParams(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=float(tf_d[0].item()),
                           dt_fast=(N_RG / 600e6) / (N_RG - 1), per_target_velocity=0,
                           samples_per_thread=dev.chunk_hint(pos0_d, ptx_d, tf_d.cpu().numpy(), prm.T_p, prm.C))
    lib, ctx = _lib.load(), _lib.context(local)
    stream = torch.cuda.current_stream(device)

    def echo_resident():
        for q0 in range(0, N_AZ, 65535):
            q1 = min(N_AZ, q0 + 65535)
            _lib.check(lib.nis_echo_accumulate(ctx, C.byref(eprm), C.c_void_p(pos0_d.data_ptr()),
                                               C.c_void_p(vel_d.data_ptr()), C.c_void_p(amp_d.data_ptr()),
                                               C.c_void_p(ptx_d.data_ptr()), C.c_void_p(0), C.c_void_p(ts_d.data_ptr()),
                                               C.c_void_p(tf_d.data_ptr()), len(sc["rcs"]), q0, q1, N_RG,
                                               C.c_void_p(raw.data_ptr()), 0, C.c_void_p(stream.cuda_stream)),
                       "nis_echo_accumulate")

    def step_resident(ev=None):
        if ev:
            ev[0].record(stream)
        echo_resident()
        if ev:
            ev[1].record(stream)
        plan.focus(raw, out=slc)
        if ev:
            ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # clocks are sampled from before the warm-up to the end of the timed region (nvidia-smi needs a
    # few 100 ms to start; the warm-up runs at least that long under the same load)
    sampler = ClockSampler(local) if rank == 0 else None
    t_w = time.perf_counter()
    n_warm = 0
    while n_warm < max(args.warmup, 3) or time.perf_counter() - t_w < 1.0:
        step_resident()
        n_warm += 1
        if n_warm % 16 == 0:
            torch.cuda.synchronize(device)
    barrier()

    # ------------------------------------------------ device-timed region (inputs resident in HBM)
    launches0 = _lib.launch_count(local)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(args.steps):
        step_resident(evs[k])
    t_end.record(stream)
    barrier()
    launches = _lib.launch_count(local) - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = t_begin.elapsed_time(t_end)
    echo_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    csa_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    nprof = min(args.steps, 64)
    stage = {k: 0.0 for k in plan.STAGES}
    for b in range(nprof):
        for k, v in plan.stage_times(b).items():
            stage[k] += v / nprof

    # ------------------------------------------------ e2e through the drop-in API (host buffers)
    args_csa = (prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, prm.t_start_fast)
    api.set_default_device(device)
    prm_api = prm

    def step_e2e():
        r, t0, _fs = api.run_physics_engine(targets, sc["pos_sat"], sc["t_vec"], params=prm_api, return_device=True)
        img, rax, cax = api.sar_focus_csa(r, *args_csa)        # complex128 numpy [N_rg, N_az] on the host
        return img

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(3):          # pinned host blocks of the caching allocator are in steady state after 2 calls
        img = step_e2e()
    del img
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        img = step_e2e()
    torch.cuda.synchronize(device)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = int(sc["pos"].nbytes + sc["rcs"].nbytes + sc["pos_sat"].nbytes + sc["t_vec"].nbytes + 8 * N_RG + 24)
    from nis_sar import hostio
    d2h_info = dict(hostio.last_transfer["d2h"] or {})
    d2h = int(d2h_info.get("pcie_bytes", img.nbytes))     # bytes that crossed PCIe; the host array is img.nbytes
    d2h_info["host_result_bytes"] = int(img.nbytes)
    del img

    # ------------------------------------------------ the three multi-GPU workloads of BASELINE.json, at this N
    plan.set_profiling(False)
    del raw, slc
    dev._plan_cache.clear()
    plan.close()
    torch.cuda.empty_cache()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak_gbs, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    multi = {}
    for key, fn in (("config4_videosar_frames", lambda: bench_config4_videosar(device, rank, world, peak_gbs, keep_pair=world == 1)),
                    ("config3_scatterer_shards", lambda: bench_config3_scatterer_shards(device, rank, world)),
                    ("config5_hrws_channel_pairs", lambda: bench_config5_hrws(device, rank, world))):
        if os.environ.get("NIS_BENCH_SKIP_MULTI"):
            break
        try:
            multi[key] = fn()
        except Exception as e:       # a failed side workload must not take the headline line with it
            multi[key] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            torch.cuda.synchronize(device)
    ati = multi.get("config4_videosar_frames")
    focused_pair = ati.pop("focused_pair", None) if isinstance(ati, dict) else None

    # ------------------------------------------------ the other BASELINE.json configs, one line each (single GPU)
    other = None
    if world == 1:
        other = {}
        # configs[0]: the reference's default ATI scene shape after the pulse shift, 7199 x 13200, two channels + GMTI
        na, nr = 7199, 13200
        pd = dev.CsaPlan(na, nr, lam=prm.Lambda, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff, r_ref=prm.R0,
                         t_start=prm.t_start_fast, device=device)
        chd = torch.view_as_complex(torch.randn((na + 1, nr, 2), device=device))
        sd1 = torch.empty((nr, na), dtype=torch.complex64, device=device)
        sd2 = torch.empty_like(sd1)
        mxd = torch.zeros(1, dtype=torch.float64, device=device)

        def default_frame():
            pd.focus(chd[1:], out=sd1, max_sq=mxd)
            pd.focus(chd[:-1], out=sd2)
            return dev.gmti_fused(sd1, sd2, max_sq=mxd, lazy=True, want=("ati_phase_masked", "dpca_mag"))
        for _ in range(2):
            default_frame()
        torch.cuda.synchronize(device)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(device)
        ea.record(cur)
        for _ in range(5):
            default_frame()
        eb.record(cur)
        torch.cuda.synchronize(device)
        dms = ea.elapsed_time(eb) / 5
        other["default_ati_scene"] = {"workload": "configs[0] shape: 2 channels x CSA 7199 x 13200 (general-size path: mixed radix "
                                                  "13200, Bluestein 7199) + fused DPCA/ATI detection",
                                      "ms_per_frame": dms, "mpixels_per_s": 2 * na * nr / (dms * 1e-3) / 1e6}
        # the whole of `python sar_ati_dcpa_sim_csa.py` at its own sizes (:184-197, :402-419): destroyer moving at 15 m/s
        # + 5000 clutter scatterers, 7200 pulses x 13200 samples, two phase centres -> shift -> CSA x2 -> DPCA/ATI
        del chd
        from nis_sar import scenes as nsc0
        ds = nsc0.ati_scene(seed=0)
        dkw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=prm.t_start_fast, fs=prm.FS,
                   n_samples=nr, device=device)
        vd = ds["vel_tx"] / np.sqrt(np.sum(ds["vel_tx"] ** 2, axis=1))[:, None]
        rawd = torch.empty((na + 1, nr), dtype=torch.complex64, device=device)

        def default_scene_full():
            for off, slc, mx in ((ds["rx_offsets"][0], sd1, mxd), (ds["rx_offsets"][1], sd2, None)):
                prx = ds["pos_tx"] + vd * off
                dev.echo_accumulate(ds["ship_pos"], ds["ship_vel"], ds["ship_rcs"], ds["pos_tx"], prx, ds["t_vec"], out=rawd, **dkw)
                dev.echo_accumulate(ds["clutter_pos"], ds["clutter_vel"], ds["clutter_rcs"], ds["pos_tx"], prx, ds["t_vec"],
                                    out=rawd, accumulate=True, **dkw)
                pd.focus(rawd[1:] if mx is not None else rawd[:-1], out=slc, max_sq=mx)
            return dev.gmti_fused(sd1, sd2, max_sq=mxd, want=("ati_phase_masked", "dpca_mag"))
        res_d = default_scene_full()
        torch.cuda.synchronize(device)
        walls = []
        for _ in range(3):                # a 0.12 s FP32-pipe-bound run: the power governor makes single samples noisy
            tw0 = time.perf_counter()
            res_d = default_scene_full()      # reads det_count / peak_idx back: host-synchronous
            torch.cuda.synchronize(device)
            walls.append(time.perf_counter() - tw0)
        tw0, tw1 = 0.0, float(np.median(walls))
        n_up = 2.0 * (len(ds["ship_rcs"]) + len(ds["clutter_rcs"])) * (na + 1) * nr
        other["default_scene_full_run"] = {
            "workload": "sar_ati_dcpa_sim_csa.py at its own sizes: 35 moving + 5000 clutter scatterers x 7200 pulses x 13200 samples x "
                        "2 phase centres (echo) -> pulse shift -> CSA 7199 x 13200 x 2 -> DPCA/ATI + detections; host scene arrays "
                        "in, detection list out",
            "s_wall": tw1 - tw0, "s_wall_runs": walls, "scatterer_sample_updates": n_up, "g_updates_per_s": n_up / (tw1 - tw0) / 1e9,
            "detections": int(res_d["det_count"])}
        pd.close()
        del rawd, sd1, sd2
        torch.cuda.empty_cache()
        # configs[2]: dense vehicle scene, 1e5 scatterers, airborne geometry (S = 2048); a 256-pulse block of the 32768
        from nis_sar import scenes as nsc
        vs = nsc.vehicle_scene(seed=0, num_pulses=256, num_scatterers=100000)
        vpr = vs["prm"]
        vkw = dict(c=vpr.C, fc=vpr.FC, k_rate=vpr.k_rate, t_p=vpr.T_p, t_start=(2 * vpr.R0 / vpr.C) - (2048 / 360e6) / 2,   # sar_vehicle_sim.py:89
                   fs=360e6, n_samples=2048, device=device)
        vargs = (vs["pos"], np.zeros(3), vs["rcs"], vs["pos_sat"], None, vs["t_vec"])
        vout = dev.echo_accumulate(*vargs, **vkw)
        torch.cuda.synchronize(device)
        ea.record(cur)
        for _ in range(3):
            dev.echo_accumulate(*vargs, out=vout, **vkw)
        eb.record(cur)
        torch.cuda.synchronize(device)
        vms = ea.elapsed_time(eb) / 3
        other["dense_vehicle_echo"] = {"workload": "configs[2]: 1e5 scatterers x 256 of 32768 pulses x 2048 samples (sar_vehicle_sim.py "
                                                   "geometry); includes the host->device copy of the scatterer arrays",
                                       "ms": vms, "g_scatterer_samples_per_s": 1e5 * 256 * 2048 / (vms * 1e-3) / 1e9,
                                       "gsamples_per_s": 256 * 2048 / (vms * 1e-3) / 1e9}
        del vout

    # ------------------------------------------------ DPCA/ATI stage alone, beside the reference's seven numpy passes
    gmti_line = None
    if world == 1 and focused_pair is not None:
        from oracle import sar_oracle as orc3
        ng = 4096
        ga, gb = focused_pair                       # the last focused two-channel frame of the VideoSAR sequence above
        mxg = torch.zeros(1, dtype=torch.float64, device=device)
        mxg.fill_(float((torch.view_as_real(ga).double() ** 2).sum(dim=-1).max()))
        cur = torch.cuda.current_stream(device)
        eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g_ms = {}
        for tag, kw in (("max_from_csa", dict(max_sq=mxg)), ("own_max_pass", {})):
            for _ in range(4):
                dev.gmti_fused(ga, gb, lazy=True, **kw)
            torch.cuda.synchronize(device)
            eg0.record(cur)
            for _ in range(20):
                dev.gmti_fused(ga, gb, lazy=True, **kw)
            eg1.record(cur)
            torch.cuda.synchronize(device)
            g_ms[tag] = eg0.elapsed_time(eg1) / 20
        rec = dev.gmti_fused(ga, gb, max_sq=mxg, lazy=True)["result_dev"].view(torch.int32).cpu()
        hs = 2048
        h1 = ga[:hs, :hs].cpu().numpy().astype(np.complex128)
        h2 = gb[:hs, :hs].cpu().numpy().astype(np.complex128)
        tg0 = time.perf_counter()
        orc3.gmti_products(h1, h2)
        g_cpu = time.perf_counter() - tg0
        gmti_line = {"workload": "4096 x 4096 focused channel pair (point scatterers over a noise floor), all seven products + detection list: "
                                 "49 B per pixel pair; max|slc1| handed over by nis_csa_focus (the path's normal case)",
                     "ms": g_ms["max_from_csa"], "ms_with_own_max_pass": g_ms["own_max_pass"], "launches": 2,
                     "detected_fraction": int(rec[0]) / float(ng * ng),
                     "mpixel_pairs_per_s": ng * ng / (g_ms["max_from_csa"] * 1e-3) / 1e6,
                     "achieved_GBps": 49.0 * ng * ng / (g_ms["max_from_csa"] * 1e-3) / 1e9,
                     "cpu_port": {"mpixel_pairs_per_s": hs * hs / g_cpu / 1e6, "cores": 1,
                                  "sample": f"the inline numpy block (sar_ati_dcpa_sim_csa.py:414-419, :447-449) on a 2048 x 2048 pair, {g_cpu:.1f} s"}}
        del ga, gb, focused_pair
        torch.cuda.empty_cache()

    # ------------------------------------------------ the reference's own GPU formulation of the echo engine, same device
    ref_gpu = None
    if world == 1:
        from nis_sar import scenes as nsc2
        from oracle import sar_oracle_torch as ot
        asc = nsc2.ati_scene(seed=0, num_pulses=8, num_clutter=5000)          # default clutter scene, 8 of 7200 pulses
        aprm = asc["prm"]
        apos = np.concatenate([asc["ship_pos"], asc["clutter_pos"]])
        arcs = np.concatenate([asc["ship_rcs"], asc["clutter_rcs"]])
        ag = aprm.as_globals()
        for _ in range(2):
            rt, _t0 = ot.echo_bistatic_torch(apos, arcs, asc["t_vec"], asc["pos_tx"], asc["vel_tx"], asc["rx_offsets"][0],
                                             asc["ship_vel"], ag, device=device)
        torch.cuda.synchronize(device)
        t0r = time.perf_counter()
        rt, _t0 = ot.echo_bistatic_torch(apos, arcs, asc["t_vec"], asc["pos_tx"], asc["vel_tx"], asc["rx_offsets"][0],
                                         asc["ship_vel"], ag, device=device)
        torch.cuda.synchronize(device)
        ref_s = time.perf_counter() - t0r
        tg_list = [{"position": p_, "rcs": r_} for p_, r_ in zip(apos, arcs)]
        mine = None
        for _ in range(3):
            t0m = time.perf_counter()
            mine, _t0 = api.run_bistatic_physics_gpu(tg_list, asc["t_vec"], asc["pos_tx"], asc["vel_tx"], asc["rx_offsets"][0],
                                                     asc["ship_vel"], params=aprm, device=device, return_device=True)
            torch.cuda.synchronize(device)
            mine_s = time.perf_counter() - t0m
        err = float(torch.linalg.vector_norm(mine.to(torch.complex128) - rt) / torch.linalg.vector_norm(rt))
        upd = float(len(arcs)) * 8 * rt.shape[1]
        asc2 = nsc2.ati_scene(seed=0, num_pulses=512, num_clutter=5000)     # enough pulses to amortise the host set-up
        for _ in range(2):
            t0m = time.perf_counter()
            m2, _t0 = api.run_bistatic_physics_gpu(tg_list, asc2["t_vec"], asc2["pos_tx"], asc2["vel_tx"],
                                                   asc2["rx_offsets"][0], asc2["ship_vel"], params=aprm, device=device,
                                                   return_device=True)
            torch.cuda.synchronize(device)
            mine512_s = time.perf_counter() - t0m
        del m2
        ref_gpu = {"workload": "default ATI clutter scene, 5035 scatterers x 8 of 7200 pulses x 13200 samples",
                   "reference_torch_eager_on_this_gpu": {"s": ref_s, "g_updates_per_s": upd / ref_s / 1e9,
                                                         "what": "oracle/sar_oracle_torch.py: the per-pulse eager torch loop of "
                                                                 "run_bistatic_physics_gpu (sar_ati_dcpa_sim_csa.py:137-178), fp64"},
                   "this_library_same_call": {"s": mine_s, "g_updates_per_s": upd / mine_s / 1e9,
                                              "what": "nis_sar.api.run_bistatic_physics_gpu incl. host list -> device arrays"},
                   "this_library_512_pulses": {"s": mine512_s, "g_updates_per_s": upd * 64 / mine512_s / 1e9},
                   "rel_l2_between_them": err}
        del rt, mine
        torch.cuda.empty_cache()

    # ------------------------------------------------ next-row N1: Range-Doppler focusing of a 4096 x 4096 frame
    rda = None
    if world == 1:
        n4 = 4096
        rprm = prm.replace(T_p=10e-6)
        rp = dev.RdaPlan(n4, n4, lam=rprm.Lambda, t_p=rprm.T_p, kr=rprm.k_rate, fs=rprm.FS, prf=rprm.PRF, vr=rprm.V_eff,
                         range_grp=rprm.R0, device=device)
        xr = torch.view_as_complex(torch.randn((n4, n4, 2), device=device))
        for _ in range(3):
            rp.focus(xr)
        torch.cuda.synchronize(device)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(device)
        ea.record(cur)
        for _ in range(20):
            rp.focus(xr)
        eb.record(cur)
        torch.cuda.synchronize(device)
        rda_ms = ea.elapsed_time(eb) / 20
        rp.close()
        del xr
        # the satellite scripts' own call (sar_satellite_sim.py:453-460): 7200 pulses x 13200 samples, T_p = 20 us -> 12001 taps
        ps, ss = 7200, 13200
        rps = dev.RdaPlan(ps, ss, lam=prm.Lambda, t_p=prm.T_p, kr=prm.k_rate, fs=prm.FS, prf=prm.PRF, vr=prm.V_eff,
                          range_grp=prm.R0, device=device)
        xsat = torch.view_as_complex(torch.randn((ps, ss, 2), device=device))
        for _ in range(2):
            rps.focus(xsat)
        torch.cuda.synchronize(device)
        ea.record(cur)
        for _ in range(5):
            rps.focus(xsat)
        eb.record(cur)
        torch.cuda.synchronize(device)
        rda_sat_ms = ea.elapsed_time(eb) / 5
        rps.close()
        del xsat
        torch.cuda.empty_cache()
        from oracle import sar_oracle as orc
        rng = np.random.default_rng(1)
        xs = rng.standard_normal((n4, 256)) + 1j * rng.standard_normal((n4, 256))
        t0c = time.perf_counter()
        orc.focus_rda(xs, rprm.Lambda, rprm.T_p, rprm.k_rate, rprm.FS, rprm.PRF, rprm.V_eff, rprm.R0)
        rda_cpu_s = time.perf_counter() - t0c
        rda_bytes = 60.0 * n4 * n4
        rda = {"workload": "4096 pulses x 4096 samples, 6001-tap matched filter (sar_focus_rda, image only)",
               "ms_per_frame": rda_ms, "mpixels_per_s": n4 * n4 / (rda_ms * 1e-3) / 1e6,
               "algorithmic_bytes_per_frame": rda_bytes, "achieved_GBps": rda_bytes / (rda_ms * 1e-3) / 1e9,
               "cpu_port": {"mpixels_per_s": xs.size / rda_cpu_s / 1e6, "cores": 1,
                            "sample": f"numpy port of sar_focus_rda on 4096 samples x 256 pulses, {rda_cpu_s:.1f} s"},
               "satellite_frame": {"workload": "sar_satellite_sim.py's own call: 7200 pulses x 13200 samples, 12001-tap matched "
                                               "filter (one zero-padded 32768-point block per pulse), image only",
                                   "ms_per_frame": rda_sat_ms, "mpixels_per_s": ps * ss / (rda_sat_ms * 1e-3) / 1e6}}

    # ------------------------------------------------ next-row N4: one VideoSAR frame of sar_batch_sim.py at its real size
    video = None
    if world == 1:
        from nis_sar import params as nparams, scenes as nscenes, targets as ntargets
        vprm = nparams.batch_spotlight_preset()
        n_p, n_pix = 2500, 512
        t_cpi = (np.arange(n_p) - (n_p - 1) / 2) / vprm.PRF
        p_cpi, v_cpi = nscenes.orbit_trajectory(vprm, t_cpi, along="x")
        ship = ntargets.generate_destroyer(center_pos=(0, 0, 0))
        l_ant = vprm.Lambda * vprm.R0 / 500.0

        def video_frame():
            r, t_st, n_sp, v_t = api.run_physics_spotlight(ship, t_cpi, p_cpi, v_cpi, 45.0, 15.0, l_ant, params=vprm,
                                                           device=device)
            dev.add_noise(r, 30.0, seed=1, ref_power="max")
            return api.tdbp_gpu(r, p_cpi, v_cpi, t_st, n_sp, v_t, t_cpi, 500.0, nx=n_pix, ny=n_pix, params=vprm,
                                device=device)             # complex128 image on the host, as the reference returns
        for _ in range(2):
            video_frame()
        torch.cuda.synchronize(device)
        t0v = time.perf_counter()
        nvf = 5
        for _ in range(nvf):
            video_frame()
        torch.cuda.synchronize(device)
        vms = (time.perf_counter() - t0v) / nvf * 1e3
        video = {"workload": "sar_batch_sim.py frame: 35-scatterer destroyer, 2500-pulse CPI x 22004 samples (12000-tap chirp), "
                             "noise at the peak power, 512 x 512 backprojection; host scene arrays in, complex128 image out",
                 "ms_per_frame_e2e": vms, "frames_per_s": 1e3 / vms,
                 "g_pixel_pulses_per_s": n_p * n_pix * n_pix / (vms * 1e-3) / 1e9}

    # ------------------------------------------------ reduce over ranks (max time)
    times = torch.tensor([total_ms, e2e_s * 1e3, echo_ms, csa_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, echo_ms_max, csa_ms_max = (float(x) for x in times.cpu())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pixels = float(N_AZ) * N_RG
    ms_per_step = total_ms / args.steps
    # share of (scatterer, pulse, sample) triples that fall inside the chirp support (exact count on a pulse subset)
    sub = np.linspace(0, N_AZ - 1, 64).astype(int)
    d = np.linalg.norm(sc["pos"][None, :, :] - sc["pos_sat"][sub][:, None, :], axis=2)
    tau = 2 * d / prm.C
    tf = dev.fast_time_axis(prm.t_start_fast, N_RG, 600e6)
    lo = np.searchsorted(tf, tau.ravel(), side="left")
    hi = np.searchsorted(tf, tau.ravel() + prm.T_p, side="right")
    in_support = float(np.mean((hi - lo) / N_RG))
    value = world * pixels / (ms_per_step * 1e-3) / 1e6
    dom = max(stage, key=stage.get)
    dom_gbs = STAGE_ALGO_BYTES_PER_PIXEL * pixels / (stage[dom] * 1e-3) / 1e9
    csa_gbs = CSA_ALGO_BYTES_PER_PIXEL * pixels / (csa_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(dom)

    cpu_v, e_rate, c_rate, cpu_dt = cpu_step_throughput(1) if world == 1 else (None, 0, 0, 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (complex64 samples; fp64 geometry and phase coefficients)", "data": "synthetic",
        "config": config_dict(args),
        "e2e": {"value": world * pixels / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "d2h_transfer": d2h_info,
                "path": "nis_sar.api.run_physics_engine(return_device=True) -> nis_sar.api.sar_focus_csa -> "
                        "complex128 numpy image on the host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"csa:{dom}", "achieved": dom_gbs, "peak": peak_gbs, "unit": "GB/s",
                     "frac": dom_gbs / peak_gbs, "traffic": traffic,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one static capture "
                                       "(profiles/traffic.json; DRAM bytes cannot be read without the profiler, and a number taken "
                                       "under ncu is never a bench value)",
                     "peak_source": peak_src,
                     "frac_of_nameplate_8TBps": dom_gbs / 8000.0,
                     "algorithmic_bytes_per_launch": STAGE_ALGO_BYTES_PER_PIXEL * pixels,
                     "launch_ms": stage[dom],
                     "csa_whole": {"achieved": csa_gbs, "frac": csa_gbs / peak_gbs,
                                   "algorithmic_bytes_per_frame": CSA_ALGO_BYTES_PER_PIXEL * pixels, "ms": csa_ms},
                     "stages_ms": stage},
        "echo": {"ms": echo_ms, "gsamples_per_s": pixels / (echo_ms * 1e-3) / 1e9,
                 "g_scatterer_samples_per_s": GRID_SIDE * GRID_SIDE * pixels / (echo_ms * 1e-3) / 1e9,
                 "in_support_fraction": in_support,
                 "g_in_support_updates_per_s": in_support * GRID_SIDE * GRID_SIDE * pixels / (echo_ms * 1e-3) / 1e9,
                 "bound": "fp32 issue / MUFU (not HBM)"},
        "csa": {"ms": csa_ms, "mpixels_per_s": pixels / (csa_ms * 1e-3) / 1e6},
        "clocks": clocks,
    }
    # Everything beyond the contract's keys is nested inside `config` / `roofline`: the driver's record keeps those dicts whole
    # and drops unknown top-level keys.
    line["config"]["multi_gpu"] = multi
    line["config"]["numa"] = numa_info
    if ati is not None and "error" not in ati:
        line["roofline"]["north_star_frame"] = {
            "what": "4096x4096 two-channel frame (CSA x2 + fused DPCA/ATI, all products) as run in config.multi_gpu.config4_videosar_frames",
            "ms_per_frame_per_gpu": ati["ms_per_frame_per_gpu"], "achieved": ati["per_gpu_achieved_GBps"],
            "frac": ati["per_gpu_frac_of_hbm_peak"], "target_frac": 0.40, "launch": ati["launch_mode"],
            "ms_per_frame_cuda_graph_replay": ati["ms_per_frame_cuda_graph_replay"],
            "frac_cuda_graph_replay": ati["frac_of_hbm_peak_cuda_graph_replay"]}
    if gmti_line is not None:
        gmti_line["frac_of_hbm_peak"] = gmti_line["achieved_GBps"] / peak_gbs
        line["roofline"]["gmti_stage"] = gmti_line
    extra = {}
    if other:
        extra.update(other)
    if ref_gpu is not None:
        extra["reference_gpu_path"] = ref_gpu
    if video is not None:
        extra["videosar_frame"] = video
    if rda is not None:
        rda["frac_of_hbm_peak"] = rda["achieved_GBps"] / peak_gbs
        extra["rda_frame"] = rda
    if extra:
        line["config"]["other_workloads"] = extra
    if cpu_v is not None:
        line["cpu_baseline"] = {
            "value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port, extrapolated",
            "sample": f"numpy port of run_physics_engine on {REF_ECHO_PULSES} of {N_AZ} pulses ({e_rate:.3g} scatterer-samples/s) + "
                      f"sar_focus_csa port on ONE {REF_CSA_N}x{REF_CSA_N} frame ({c_rate / 1e6:.3g} Mpixel/s); per-pixel costs "
                      f"added and scaled to the {N_AZ}x{N_RG} step (a model of the full step, {cpu_dt:.1f} s of CPU work)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
