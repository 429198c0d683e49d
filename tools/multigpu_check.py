#!/usr/bin/env python
"""Multi-GPU correctness + timing of the sharded drivers in nis_sar.dist (NCCL over NVLink).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

Checks, with N ranks against a one-GPU computation on rank 0:
  1. scatterer-sharded echo + sum all-reduce (config 3 style)        -> rel-L2 vs unsharded
  2. pulse-block echo + all-gather                                    -> bit-equal vs unsharded
  3. frame-parallel CSA (config 4 style)                              -> frames identical to a local recompute
  4. one receive channel per rank, ring exchange, DPCA/ATI per pair   -> indices equal to a local recompute
  5. VideoSAR frames (spotlight echo -> TDBP), round robin            -> frames identical to a local recompute
  1b / 4b: the two exchange steps with the transfer fused into the kernels (peer-mapped HBM over NVLink)
  6. HRWS-N fast-mover scene (config 5): echo -> CSA per channel per rank, pairs over NVLink -> equal to a local
     recompute; ATI phase of the mover and clutter cancellation per pair
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

import numpy as np
import torch
import torch.distributed as dist

from nis_sar import device as dev, dist as nd, params, scenes


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    out = {"world": world}

    # ---- 1/2: echo sharding on the airborne dense-vehicle geometry
    sc = scenes.vehicle_scene(seed=3, num_pulses=512, num_scatterers=4096)
    prm = sc["prm"]
    t0 = (2 * prm.R0 / prm.C) - (2048 / 360e6) / 2
    kw = dict(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=t0, fs=360e6, n_samples=2048, device=device)
    full = dev.echo_accumulate(sc["pos"], np.zeros(3), sc["rcs"], sc["pos_sat"], None, sc["t_vec"], **kw)

    def shard(a, b):
        return dev.echo_accumulate(sc["pos"][a:b], np.zeros(3), sc["rcs"][a:b], sc["pos_sat"], None, sc["t_vec"], **kw)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    summed = nd.echo_scatterer_shards(shard, len(sc["rcs"]))      # warm-up (NCCL communicator set-up)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    summed = nd.echo_scatterer_shards(shard, len(sc["rcs"]))
    ev1.record()
    torch.cuda.synchronize()
    out["scatterer_shards_rel_l2"] = float(torch.linalg.vector_norm(summed - full) / torch.linalg.vector_norm(full))
    out["scatterer_shards_ms"] = nd.max_over_ranks(ev0.elapsed_time(ev1), device)

    # ---- 1b: the same reduction fused into the synthesis kernel: every rank adds its scatterers' echo straight into the
    #          owner's pulse block over NVLink (peer-mapped buffers, RED.ADD), no partial echo, no collective
    shared = nd.SharedBuffer(tuple(full.shape), torch.complex64, device=device)

    def shard_into(a, b, p0, p1, dst):
        dev.echo_accumulate(sc["pos"][a:b], np.zeros(3), sc["rcs"][a:b], sc["pos_sat"], None, sc["t_vec"], out=dst,
                            pulse_range=(p0, p1), accumulate="atomic", **kw)
    try:
        own = nd.echo_scatterer_shards_p2p(shard_into, len(sc["rcs"]), shared)      # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        own = nd.echo_scatterer_shards_p2p(shard_into, len(sc["rcs"]), shared)
        ev1.record()
        torch.cuda.synchronize()
        blk = shared.local[own[0]:own[1]]
        ref = full[own[0]:own[1]]
        err = torch.linalg.vector_norm(blk - ref) / torch.linalg.vector_norm(ref)
        errs = torch.tensor([float(err)], device=device)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        out["scatterer_shards_p2p_rel_l2"] = float(errs.item())
        out["scatterer_shards_p2p_ms"] = nd.max_over_ranks(ev0.elapsed_time(ev1), device)
    except Exception as e:   # no peer path on this box: report, the NCCL route above is the fallback
        out["scatterer_shards_p2p_error"] = str(e)[:200]

    raw = torch.zeros_like(full)

    def fill(p0, p1, buf):
        dev.echo_accumulate(sc["pos"], np.zeros(3), sc["rcs"], sc["pos_sat"], None, sc["t_vec"], out=buf,
                            pulse_range=(p0, p1), **kw)
    nd.echo_pulse_blocks(fill, raw, gather=True)
    out["pulse_blocks_equal"] = bool(torch.equal(raw, full))

    # ---- 3: frame-parallel CSA
    sp = params.spaceborne_preset()
    n = 1024
    plan = dev.cached_plan(n, n, lam=sp.Lambda, kr=sp.k_rate, fs=sp.FS, prf=sp.PRF, vr=sp.V_eff, r_ref=sp.R0,
                           t_start=sp.t_start_fast, device=device)

    def frame(f):
        g = torch.Generator(device=device).manual_seed(1000 + f)
        return torch.view_as_complex(torch.randn((n, n, 2), generator=g, device=device))
    mine = nd.focus_frames(lambda f: plan.focus(frame(f)).clone(), 8)
    out["frames_owned"] = sorted(mine)
    chk = all(torch.equal(v, plan.focus(frame(f))) for f, v in mine.items())
    flag = torch.tensor([1 if chk else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["frames_ok"] = bool(flag.item())

    # ---- 4: one receive channel per rank, neighbour exchange, fused DPCA/ATI per adjacent pair
    chan = plan.focus(frame(100 + rank)).clone()
    pair = nd.pair_products(chan, lambda a, b: dev.gmti_fused(a, b, want=("ati_phase_masked",)))
    ok = 1
    if rank < world - 1:
        ref = dev.gmti_fused(chan, plan.focus(frame(100 + rank + 1)).clone(), want=("ati_phase_masked",))
        ok = int(torch.equal(pair["det_idx"], ref["det_idx"]) and pair["peak_idx"] == ref["peak_idx"])
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["channel_pairs_ok"] = bool(flag.item())

    # ---- 4b: pairing at 4096^2, staging copy (isend / irecv) vs the DPCA/ATI kernel reading the neighbour's HBM over NVLink
    n4 = 4096
    p4 = dev.cached_plan(n4, n4, lam=sp.Lambda, kr=sp.k_rate, fs=sp.FS, prf=sp.PRF, vr=sp.V_eff, r_ref=sp.R0,
                         t_start=sp.t_start_fast, device=device)
    g4 = torch.Generator(device=device).manual_seed(500 + rank)
    chan4 = p4.focus(torch.view_as_complex(torch.randn((n4, n4, 2), generator=g4, device=device))).clone()
    prod = lambda a, b: dev.gmti_fused(a, b, lazy=True)       # all products, no host read-back
    try:
        sh4 = nd.SharedBuffer((n4, n4), torch.complex64, device=device)
        sh4.local.copy_(chan4)
        for fn_name, fn in (("copy", lambda: nd.pair_products(chan4, prod)), ("p2p", lambda: nd.pair_products_p2p(sh4, prod))):
            r0 = fn()
            torch.cuda.synchronize()
            dist.barrier()
            ev0.record()
            for _ in range(5):
                r0 = fn()
            ev1.record()
            torch.cuda.synchronize()
            out[f"pair4096_{fn_name}_ms"] = nd.max_over_ranks(ev0.elapsed_time(ev1) / 5, device)
            if r0 is not None:
                out.setdefault("_pair_results", {})[fn_name] = (r0["det_idx_raw"][:1000].clone(), r0["result_dev"].clone())
        ok = 1
        pr = out.pop("_pair_results", None)
        if pr:
            ok = int(torch.equal(pr["copy"][0], pr["p2p"][0]) and torch.equal(pr["copy"][1], pr["p2p"][1]))
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["pair4096_p2p_equals_copy"] = bool(flag.item())
    except Exception as e:
        out.pop("_pair_results", None)
        out["pair4096_p2p_error"] = str(e)[:200]
    # ---- 5: VideoSAR frames (spotlight echo -> backprojection), round robin over ranks, timed
    from nis_sar import video, targets as tg
    vp = params.batch_spotlight_preset(fs=60e6, bw=50e6, t_p=2e-6)
    total, step, cpi, nfr = 1000, 100, 400, 7
    t_all = np.linspace(-0.1, 0.1, total)
    pos_all, vel_all = scenes.orbit_trajectory(vp, t_all, along="x")
    base = tg.generate_destroyer(center_pos=(0, 0, 0))
    kwv = dict(heading_deg=45.0, speed=15.0, l_ant=vp.Lambda * vp.R0 / 500.0, scene_size=500.0, step_pulses=step,
               cpi_pulses=cpi, num_frames=nfr, nx=128, ny=128, params=vp, device=device, return_device=True)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    mine_v = video.render_frames(base, t_all, pos_all, vel_all, **kwv)
    ev1.record()
    torch.cuda.synchronize()
    out["video_frames_owned"] = sorted(mine_v)
    out["video_ms_all_frames"] = nd.max_over_ranks(ev0.elapsed_time(ev1), device)
    chk = all(torch.equal(v, video.render_frames(base, t_all, pos_all, vel_all, frames=[f], **kwv)[f]) for f, v in mine_v.items())
    flag = torch.tensor([1 if chk else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["video_frames_ok"] = bool(flag.item())
    counts = torch.tensor([len(mine_v)], device=device)
    dist.all_reduce(counts)
    out["video_frames_total"] = int(counts.item())
    # ---- 6: HRWS-N fast-mover scene (config 5): channel k on rank k (echo -> one-pulse DPCA shift -> CSA), pair (k, k+1)
    #         formed by the DPCA/ATI kernel reading rank k+1's image over NVLink
    from nis_sar import api
    hp = params.spaceborne_preset(fs=100e6, bw=80e6).replace(n_samples=2048, window_s=2048 / 100e6)
    hs = scenes.hrws_scene(world, seed=17, num_pulses=1025, num_clutter=400, prm=hp, t_int=None)
    ship, clut = (hs["ship_pos"], hs["ship_rcs"]), (hs["clutter_pos"], hs["clutter_rcs"])

    def channel(k):
        a, _ = api.run_bistatic_physics_gpu(tg.arrays_to_targets(*ship), hs["t_vec"], hs["pos_tx"], hs["vel_tx"],
                                            hs["rx_offsets"][k], hs["ship_vel"], params=hp, device=device, return_device=True)
        b, _ = api.run_bistatic_physics_gpu(tg.arrays_to_targets(*clut), hs["t_vec"], hs["pos_tx"], hs["vel_tx"],
                                            hs["rx_offsets"][k], hs["clutter_vel"], params=hp, device=device, return_device=True)
        return a + b

    def focus(raw):
        return api.sar_focus_csa(raw, hp.Lambda, hp.T_p, hp.k_rate, hp.FS, hp.PRF, hp.V_eff, hp.R0, hp.t_start_fast,
                                 device=device, return_device=True)[0].clone()
    try:
        raw_k = channel(rank)
        lead = focus(raw_k[1:].contiguous())                   # slc1 of pair (k, k+1)
        shh = nd.SharedBuffer(tuple(lead.shape), torch.complex64, device=device)
        shh.local.copy_(focus(raw_k[:-1].contiguous()))        # slc2 of pair (k-1, k), read by rank k-1
        prod_h = lambda a, b: dev.gmti_fused(a, b, want=("ati_phase", "ati_phase_masked", "dpca_mag"))
        res = nd.pair_products_p2p(shh, prod_h, first=lead)
        ok, stats = 1, torch.zeros(5, device=device, dtype=torch.float64)
        if res is not None:
            ref = prod_h(lead, focus(channel(rank + 1)[:-1].contiguous()))
            ok = int(torch.equal(res["det_idx"], ref["det_idx"]) and res["peak_idx"] == ref["peak_idx"]
                     and torch.equal(res["dpca_mag"], ref["dpca_mag"]))
            mag = lead.abs()
            dp = res["dpca_mag"]
            mover = int(torch.argmax(dp))                       # the ship survives the subtraction
            clutter_px = int(torch.argmax(torch.where(dp < 0.05 * dp.max(), mag, torch.zeros_like(mag))))
            stats = torch.tensor([float(res["ati_phase"].flatten()[mover]),
                                  20 * np.log10(float(dp.flatten()[clutter_px]) / float(mag.flatten()[clutter_px])),
                                  float(len(res["det_idx"])), float(dp.flatten()[mover] / mag.flatten()[mover]),
                                  float(mag.flatten()[mover] / mag.max())], device=device, dtype=torch.float64)
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["hrws_pairs_equal_local_recompute"] = bool(flag.item())
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        out["hrws_mover_ati_phase_rad_per_pair"] = [round(float(s[0]), 4) for s in allst[:-1]]
        out["hrws_clutter_cancellation_db_per_pair"] = [round(float(s[1]), 1) for s in allst[:-1]]
        out["hrws_detections_per_pair"] = [int(s[2]) for s in allst[:-1]]
        out["hrws_mover_dpca_over_slc_and_rel_mag"] = [[round(float(s[3]), 3), round(float(s[4]), 4)] for s in allst[:-1]]
    except Exception as e:
        out["hrws_error"] = str(e)[:300]
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
