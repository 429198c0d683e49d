"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the reference itself.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Every fixture is the output of the reference's *own* code (AST-extracted, unmodified --
see oracle/ref_extract.py) on seeded inputs that are stored next to it, so the tests can
run anywhere without the reference tree.  Large outputs are stored decimated together
with whole-array checksums (sum, sum of squares) so fixtures stay small.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nis-sar-amtigmti-video_b200"))

from oracle import ref_extract  # noqa: E402
from nis_sar import scenes, params  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _as_targets(pos, rcs):
    return [{"position": np.array(p), "rcs": float(r)} for p, r in zip(pos, rcs)]


def _digest(a, step):
    a = np.asarray(a)
    flat = a.ravel()
    return {"dec": flat[::step].copy(), "sum": np.sum(flat), "sumsq": np.sum(np.abs(flat) ** 2),
            "shape": np.array(a.shape)}


def golden_vehicle_targets():
    gens = ref_extract.vehicle_target_generators()
    out = {}
    center = (12.5, -40.0, 3.0)
    for name, fn in gens.items():
        t = fn(center_pos=center)
        out[f"{name}_pos"] = np.array([x["position"] for x in t], dtype=np.float64)
        out[f"{name}_rcs"] = np.array([x["rcs"] for x in t], dtype=np.float64)
    out["center"] = np.array(center)
    np.savez_compressed(os.path.join(OUT, "vehicle_targets.npz"), **out)


def golden_echo():
    # (1) bistatic engine at reduced sample rate: full arrays.
    prm = params.spaceborne_preset(fs=60e6, bw=50e6)
    sc = scenes.ati_scene(seed=11, num_pulses=8, num_clutter=20, prm=prm)
    bist, _ = ref_extract.ati_csa_functions(prm.as_globals())
    raw_ship, t0 = bist(_as_targets(sc["ship_pos"], sc["ship_rcs"]), sc["t_vec"], sc["pos_tx"], sc["vel_tx"],
                        sc["rx_offsets"][0], sc["ship_vel"])
    raw_clut, _ = bist(_as_targets(sc["clutter_pos"], sc["clutter_rcs"]), sc["t_vec"], sc["pos_tx"], sc["vel_tx"],
                       sc["rx_offsets"][1], sc["clutter_vel"])
    np.savez_compressed(os.path.join(OUT, "echo_bistatic_fs60.npz"),
                        fs=60e6, bw=50e6, seed=11, num_pulses=8, num_clutter=20,
                        raw_ship_rx1=raw_ship, raw_clutter_rx2=raw_clut, t_start_fast=t0)

    # (2) bistatic engine at the true 600 MHz parameters: decimated + checksums + support.
    prm = params.spaceborne_preset()
    sc = scenes.ati_scene(seed=12, num_pulses=3, num_clutter=30, prm=prm)
    bist, _ = ref_extract.ati_csa_functions(prm.as_globals())
    pos = np.concatenate([sc["ship_pos"], sc["clutter_pos"]])
    rcs = np.concatenate([sc["ship_rcs"], sc["clutter_rcs"]])
    raw, t0 = bist(_as_targets(pos, rcs), sc["t_vec"], sc["pos_tx"], sc["vel_tx"], sc["rx_offsets"][0], sc["ship_vel"])
    d = _digest(raw, 7)
    nz = raw != 0
    np.savez_compressed(os.path.join(OUT, "echo_bistatic_fs600.npz"), seed=12, num_pulses=3, num_clutter=30,
                        dec=d["dec"], step=7, sum=d["sum"], sumsq=d["sumsq"], shape=d["shape"], t_start_fast=t0,
                        first_nz=np.argmax(nz, axis=1), last_nz=raw.shape[1] - 1 - np.argmax(nz[:, ::-1], axis=1))

    # (3) monostatic engines (S is hard-wired to 13200 / 2048 inside them).
    g = prm.as_globals()
    sat = scenes.stripmap_scene(num_pulses=2, num_samples=13200, n_side=3, half_extent=400.0)
    eng = ref_extract.satellite_engine(g)
    raw, t0, fs = eng(_as_targets(sat["pos"], sat["rcs"]), sat["pos_sat"], sat["t_vec"])
    d = _digest(raw, 5)
    np.savez_compressed(os.path.join(OUT, "echo_satellite.npz"), dec=d["dec"], step=5, sum=d["sum"],
                        sumsq=d["sumsq"], shape=d["shape"], t_start_fast=t0, fs=fs)

    mov = ref_extract.moving_engine(g)
    vel = [4.0, -9.0, 0.0]
    raw, t0, fs = mov(_as_targets(sc["ship_pos"], sc["ship_rcs"]), sat["t_vec"], sat["pos_sat"], vel)
    d = _digest(raw, 5)
    np.savez_compressed(os.path.join(OUT, "echo_moving.npz"), dec=d["dec"], step=5, sum=d["sum"],
                        sumsq=d["sumsq"], shape=d["shape"], t_start_fast=t0, fs=fs, vel=np.array(vel))

    vp = params.airborne_vehicle_preset()
    veh = scenes.vehicle_scene(seed=5, num_pulses=4, num_scatterers=60)
    # spread the 4 pulses over the real aperture so the range history actually changes
    t_vec = np.linspace(-8.0, 8.0, 4)
    pos_plat, _ = scenes.straight_trajectory(vp, t_vec)
    ve = ref_extract.vehicle_engine(vp.as_globals())
    raw = ve(_as_targets(veh["pos"], veh["rcs"]), t_vec, pos_plat, 500e-6, vp.T_p, vp.FC, vp.BW)
    np.savez_compressed(os.path.join(OUT, "echo_vehicle.npz"), raw=raw, t_vec=t_vec, seed=5, num_scatterers=60)


def golden_csa():
    prm = params.spaceborne_preset()
    _, csa = ref_extract.ati_csa_functions(prm.as_globals())
    rng = np.random.default_rng(2024)
    out = {}
    for tag, (n_az, n_rg) in {"p2": (64, 128), "odd": (45, 88), "prime": (31, 101)}.items():
        x = (rng.standard_normal((n_az, n_rg)) + 1j * rng.standard_normal((n_az, n_rg))).astype(np.complex64)
        img, rax, cax = csa(x.astype(np.complex128), prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF,
                            prm.V_eff, prm.R0, prm.t_start_fast)
        assert img.shape == (n_rg, n_az) and img.flags.f_contiguous
        out[f"{tag}_in"] = x
        out[f"{tag}_img"] = np.ascontiguousarray(img)
        out[f"{tag}_rax"] = rax
        out[f"{tag}_cax"] = cax
    np.savez_compressed(os.path.join(OUT, "csa_random.npz"), **out)


def golden_chain():
    """Reduced default scene end to end through the reference: 2-channel echo, pulse shift,
    CSA x2, inline ATI/DPCA products.  Stored as digests."""
    prm = params.spaceborne_preset()
    P = 64
    sc = scenes.ati_scene(seed=7, num_pulses=P, num_clutter=40, prm=prm)
    bist, csa = ref_extract.ati_csa_functions(prm.as_globals())
    ship = _as_targets(sc["ship_pos"], sc["ship_rcs"])
    clut = _as_targets(sc["clutter_pos"], sc["clutter_rcs"])
    raws = []
    for off in sc["rx_offsets"]:
        a, t0 = bist(ship, sc["t_vec"], sc["pos_tx"], sc["vel_tx"], off, sc["ship_vel"])
        b, _ = bist(clut, sc["t_vec"], sc["pos_tx"], sc["vel_tx"], off, sc["clutter_vel"])
        raws.append(a + b)
    rx1, rx2 = raws[0][1:, :], raws[1][:-1, :]           # sar_ati_dcpa_sim_csa.py:402-403
    slc1, rax, cax = csa(rx1, prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, t0)
    slc2, _, _ = csa(rx2, prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, t0)
    prod = ref_extract.ati_inline_products(slc1, slc2)
    step = 97
    mag = prod["slc1_mag"]
    thr = mag.max() * 0.05
    np.savez_compressed(
        os.path.join(OUT, "chain_ati_p64.npz"), seed=7, num_pulses=P, num_clutter=40, step=step,
        slc1_dec=np.ascontiguousarray(slc1).ravel()[::step], slc2_dec=np.ascontiguousarray(slc2).ravel()[::step],
        slc1_sumsq=np.sum(np.abs(slc1) ** 2), slc2_sumsq=np.sum(np.abs(slc2) ** 2),
        interf_dec=np.ascontiguousarray(prod["ati_interf"]).ravel()[::step],
        phase_dec=np.ascontiguousarray(prod["ati_phase"]).ravel()[::step],
        dpca_mag_dec=np.ascontiguousarray(prod["dpca_mag"]).ravel()[::step],
        det_idx=np.flatnonzero(prod["mag_mask"]), peak_idx=int(np.argmax(mag)),
        phase_at_det=prod["ati_phase_masked"][prod["mag_mask"]],
        margin=float(np.min(np.abs(mag - thr)) / thr), shape=np.array(slc1.shape),
        rax=rax, cax=cax, t_start_fast=t0,
        f_contiguous=bool(slc1.flags.f_contiguous))


def golden_rda():
    """sar_focus_rda of the three simulators on seeded random phase histories: a power-of-two case, a
    smooth non-power-of-two case and an odd x odd case (exercises the (n-1)/2 axis branches)."""
    prm = params.spaceborne_preset(fs=60e6, bw=50e6).replace(T_p=1e-6)
    fns = ref_extract.rda_functions()
    rng = np.random.default_rng(77)
    out = {"t_p": prm.T_p, "kr": prm.k_rate, "fs": prm.FS, "prf": prm.PRF, "vr": prm.V_eff, "r0": prm.R0, "lam": prm.Lambda}
    for tag, (nr, npul) in {"p2": (256, 128), "smooth": (200, 96), "odd": (131, 45)}.items():
        x = (rng.standard_normal((nr, npul)) + 1j * rng.standard_normal((nr, npul))).astype(np.complex64)
        args = (x.astype(np.complex128), prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0)
        img, rax, cax, rc, rd, rcmc, filt, dop = fns["vehicle"](*args)
        img7 = fns["satellite"](*args)
        img3 = fns["moving"](*args)
        assert len(img7) == 7 and len(img3) == 3
        assert np.array_equal(img7[0], img) and np.array_equal(img3[0], img) and np.array_equal(img7[5], rcmc)
        out[f"{tag}_in"] = x
        out[f"{tag}_img"] = img
        out[f"{tag}_rax"] = rax
        out[f"{tag}_cax"] = cax
        # intermediates: every 3rd range sample x every 2nd pulse (the odd case in full)
        sl = (slice(None), slice(None)) if tag == "odd" else (slice(None, None, 3), slice(None, None, 2))
        out[f"{tag}_rc"] = rc[sl].astype(np.complex64)
        out[f"{tag}_rd"] = rd[sl].astype(np.complex64)
        out[f"{tag}_rcmc"] = rcmc[sl].astype(np.complex64)
        out[f"{tag}_filt"] = filt[sl].astype(np.complex64)
        out[f"{tag}_dop"] = dop
    np.savez_compressed(os.path.join(OUT, "rda_random.npz"), **out)


def golden_noise():
    """calculate_snr_db of the satellite and airborne scripts, and add_ocean_noise after np.random.seed (the legacy
    global generator the reference draws from) on a small echo -- pins the oracle's draw order bit for bit."""
    out = {}
    rng = np.random.default_rng(3)
    raw = (rng.standard_normal((24, 40)) + 1j * rng.standard_normal((24, 40))) * 3.0
    for key, fname in (("satellite", "sar_satellite_sim.py"), ("vehicle", "sar_vehicle_sim.py")):
        snr_fn, noise_fn = ref_extract.noise_functions(fname)
        args = np.array([[6.1e5, 5.0e4, 0.031, 5.0e8, 1.2], [2.8e4, 10.0, 0.03, 3.0e8, 16.384]])
        out[f"{key}_snr_args"] = args
        out[f"{key}_snr"] = np.array([snr_fn(*a) for a in args])
        for nu in (1.0, 0.5, 3.7):
            np.random.seed(11)
            out[f"{key}_noisy_nu{nu}"] = noise_fn(raw, 17.0, 10.0, nu)
    out["raw"] = raw
    raw_fn = ref_extract.batch_snr_function()
    bargs = np.array([[6.1e5, 5000.0, 0.031, 5.0e8, 37.9], [6.1e5, 5.0, 0.031, 5.0e8, 9.5]])
    out["batch_snr_args"] = bargs
    out["batch_snr"] = np.array([raw_fn(*a) for a in bargs])
    np.savez_compressed(os.path.join(OUT, "noise.npz"), **out)


def golden_viewer():
    """SARData (sar_ati_dcpa_viewer_csa.py:35-55) on a small seeded channel pair, before and after a balance."""
    cls = ref_extract.viewer_sardata_class()
    rng = np.random.default_rng(21)
    s1 = rng.standard_normal((20, 30)) + 1j * rng.standard_normal((20, 30))
    s2 = s1 * np.exp(0.4j) + 0.2 * (rng.standard_normal((20, 30)) + 1j * rng.standard_normal((20, 30)))
    sar = cls(s1, s2)
    out = {"s1": s1, "s2": s2}
    for m, v in sar.prods.items():
        out["cal0_" + m] = v
    sar.cal_phase = np.angle(np.mean(s1 * np.conj(s2)))
    sar.compute_all()
    out["cal_phase"] = sar.cal_phase
    for m, v in sar.prods.items():
        out["cal1_" + m] = v
    np.savez_compressed(os.path.join(OUT, "viewer.npz"), **out)


def batch_globals(fs=60e6, bw=50e6, t_p=2e-6):
    prm = params.spaceborne_preset()
    return {"C": prm.C, "R0": prm.R0, "FC": 9.65e9, "T_P": t_p, "K_RATE": bw / t_p, "FS": fs, "Lambda": prm.C / 9.65e9}


def golden_spotlight_tdbp():
    """run_physics_spotlight + tdbp_gpu of sar_batch_sim.py on the CPU (torch): a 48-pulse CPI of the destroyer moving at
    15 m/s on heading 45 deg, backprojected on a 24 x 24 grid with and without target-motion focusing (mBP / StdBP)."""
    from nis_sar import targets as tg
    g = batch_globals()
    spot, tdbp = ref_extract.batch_functions(g)
    prm = params.spaceborne_preset(prf=5000.0)
    n_p = 48
    t_vec = (np.arange(n_p) - (n_p - 1) / 2) / 5000.0 + 0.03
    pos_sat, vel_sat = scenes.orbit_trajectory(prm, t_vec, along="x")
    base = tg.generate_destroyer(center_pos=(0, 0, 0))[::2]
    l_ant = g["Lambda"] * g["R0"] / 500.0
    raw, t_st, n_sp, v_tgt = spot(base, t_vec, pos_sat, vel_sat, heading_deg=45, speed=15.0, l_ant=l_ant)
    out = {"g_keys": np.array(list(g.keys())), "g_vals": np.array(list(g.values())), "t_vec": t_vec, "pos_sat": pos_sat,
           "vel_sat": vel_sat, "l_ant": l_ant, "raw": raw.numpy().astype(np.complex64), "t_start": t_st, "n_samples": n_sp,
           "v_tgt": v_tgt, "pos0": np.array([t["position"] for t in base], dtype=float),
           "rcs": np.array([t["rcs"] for t in base], dtype=float)}
    for tag, vf in (("mbp", v_tgt), ("stdbp", np.zeros(3))):
        out["img_" + tag] = tdbp(raw, pos_sat, vel_sat, t_st, n_sp, vel_focus=vf, t_pulses=t_vec, scene_size=500.0, nx=24, ny=24)
    np.savez_compressed(os.path.join(OUT, "spotlight_tdbp.npz"), **out)


def golden_csa_large():
    """The reference's own sar_focus_csa at the two full sizes the GPU tests could not reach with an in-test oracle:
    8192 x 8192 (BASELINE.json configs[1], the bench frame) and 7199 x 13200 (the default scene after the pulse shift,
    sar_ati_dcpa_sim_csa.py:402-411).  ~1 min and ~25 GB each; stored as digests (oracle/inputs.image_digest)."""
    from oracle import inputs
    prm = params.spaceborne_preset()
    out = {}
    for tag, n_az, n_rg, seed in (("p8192", 8192, 8192, 8192), ("default", 7199, 13200, 7199)):
        prm_t = prm if tag == "default" else prm.replace(n_samples=n_rg, window_s=n_rg / 600e6)
        _, csa = ref_extract.ati_csa_functions(prm_t.as_globals())
        x = inputs.large_csa_input(n_az, n_rg, seed)
        img, rax, cax = csa(x.astype(np.complex128), prm_t.Lambda, prm_t.T_p, prm_t.k_rate, prm_t.FS, prm_t.PRF,
                            prm_t.V_eff, prm_t.R0, prm_t.t_start_fast)
        del x
        assert img.shape == (n_rg, n_az)
        d = inputs.image_digest(img, 4099)
        del img
        for k, v in d.items():
            out[f"{tag}_{k}"] = v.astype(np.complex64) if k in ("dec", "rows", "cols") else v
        out[f"{tag}_seed"] = seed
        out[f"{tag}_rax"] = rax
        out[f"{tag}_cax"] = cax
        out[f"{tag}_t_start"] = prm_t.t_start_fast
        print(tag, "sumsq", d["sumsq"], flush=True)
    np.savez_compressed(os.path.join(OUT, "csa_large.npz"), **out)


def golden_chain_4096():
    """The north-star frame at its real size through the UNMODIFIED reference: two-channel echo of the destroyer + 40 clutter
    scatterers (run_bistatic_physics_gpu x 4, 4097 pulses x 4096 samples -- FS chosen so that int(22e-6 FS) = 4096), pulse
    shift, sar_focus_csa x 2 at 4096 x 4096, the inline ATI / DPCA block.  ~3 min here; stored as digests."""
    from oracle import inputs
    fs = 4096.5 / 22e-6
    prm = params.spaceborne_preset(fs=fs, bw=150e6)
    assert int(22e-6 * prm.FS) == 4096
    P = 4097
    sc = scenes.ati_scene(seed=41, num_pulses=P, num_clutter=40, prm=prm, t_int=None)
    bist, csa = ref_extract.ati_csa_functions(prm.as_globals())
    ship = _as_targets(sc["ship_pos"], sc["ship_rcs"])
    clut = _as_targets(sc["clutter_pos"], sc["clutter_rcs"])
    raws = []
    for off in sc["rx_offsets"]:
        a, t0 = bist(ship, sc["t_vec"], sc["pos_tx"], sc["vel_tx"], off, sc["ship_vel"])
        b, _ = bist(clut, sc["t_vec"], sc["pos_tx"], sc["vel_tx"], off, sc["clutter_vel"])
        raws.append(a + b)
        print("echo channel done", raws[-1].shape, flush=True)
    rx1, rx2 = raws[0][1:, :], raws[1][:-1, :]
    slc1, rax, cax = csa(rx1, prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, t0)
    slc2, _, _ = csa(rx2, prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, t0)
    prod = ref_extract.ati_inline_products(slc1, slc2)
    mag = prod["slc1_mag"]
    thr = mag.max() * 0.05
    out = {"seed": 41, "num_pulses": P, "num_clutter": 40, "fs": fs, "bw": 150e6, "t_start_fast": t0, "rax": rax, "cax": cax}
    d = inputs.image_digest(np.ascontiguousarray(raws[0]), 4099)
    out.update({f"raw1_{k}": (v.astype(np.complex64) if k in ("dec", "rows", "cols") else v) for k, v in d.items()})
    for tag, img in (("slc1", slc1), ("slc2", slc2), ("interf", prod["ati_interf"]), ("dpca", prod["dpca_diff"])):
        d = inputs.image_digest(np.ascontiguousarray(img), 4099)
        out.update({f"{tag}_{k}": (v.astype(np.complex64) if k in ("dec", "rows", "cols") else v) for k, v in d.items()})
    out["det_idx"] = np.flatnonzero(prod["mag_mask"]).astype(np.int32)
    out["peak_idx"] = int(np.argmax(mag))
    out["phase_at_det"] = prod["ati_phase_masked"][prod["mag_mask"]].astype(np.float32)
    out["margin"] = float(np.min(np.abs(mag - thr)) / thr)
    out["dpca_peak_idx"] = int(np.argmax(prod["dpca_mag"]))
    print("detections", len(out["det_idx"]), "margin", out["margin"], flush=True)
    np.savez_compressed(os.path.join(OUT, "chain_ati_4096.npz"), **out)


if __name__ == "__main__":
    if not ref_extract.reference_available():
        sys.exit("reference tree not found: fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "large":      # the two full-size CSA frames only (minutes, ~30 GB)
        golden_csa_large()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "chain4096":  # the north-star two-channel frame at full size (minutes)
        golden_chain_4096()
        sys.exit(0)
    golden_vehicle_targets()
    golden_echo()
    golden_csa()
    golden_chain()
    golden_rda()
    golden_noise()
    golden_viewer()
    golden_spotlight_tdbp()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
