"""Drop-in replacements for the reference's hot-path entry points.

Same names, positional signatures, return tuples, dtypes (complex128 / float64) and shapes as

    run_bistatic_physics_gpu   sar_ati_dcpa_sim_csa.py:106-181
    sar_focus_csa              sar_ati_dcpa_sim_csa.py:202-396
    run_physics_engine         sar_satellite_sim.py:211-305
    run_moving_physics         sar_satellite_moving_sim.py:111-159
    run_custom_physics         sar_vehicle_sim.py:83-126

plus ``gmti_products`` for the inline ATI/DPCA block (sar_ati_dcpa_sim_csa.py:414-419, :447-449).
The reference functions read radar constants from their module's globals; here they come from a
``RadarParams`` -- either the module default (``set_default_params``) or, after
``install(module_globals)``, from the patched module's own ``C, R0, FC, BW, T_p, FS`` at call time.
Everything is computed by the CUDA library; host arrays are only copied in and out.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev
from . import hostio
from .params import RadarParams, spaceborne_preset
from .targets import targets_to_arrays

_default_params: RadarParams | None = None
_default_device = "cuda"


def set_default_params(prm: RadarParams | None):
    global _default_params
    _default_params = prm


def set_default_device(device):
    global _default_device
    _default_device = device


def _params(prm):
    if prm is not None:
        return prm
    return _default_params if _default_params is not None else spaceborne_preset()


def _to_host_c128(t: torch.Tensor, out=None) -> np.ndarray:
    """Device complex64 -> host complex128 numpy (``nis_sar.hostio``: widened on the device chunk by chunk, every chunk's
    D2H copy overlapping the next chunk's widening, into ``out`` or a pinned block of the per-device pool)."""
    return hostio.to_host_c128(t, out=out)


def _to_host_c128_fortran(t: torch.Tensor, out=None) -> np.ndarray:
    """The same values as an F-contiguous [rows, cols] array: the reference's ``img.T`` (sar_ati_dcpa_sim_csa.py:396) is a
    view of a C-ordered [n_az, n_rg] buffer, and callers see that in ``.flags`` / ``.strides`` and in how ``np.savez``
    (:457-461) lays the file out.  The corner turn back runs on the device (one 16 B/pixel pass, hidden behind PCIe)."""
    return hostio.to_host_c128(dev.transpose_c32(t.contiguous()), out=None if out is None else out.T).T


def _upload_c32(a, device) -> torch.Tensor:
    """numpy complex (any layout) or torch tensor -> contiguous complex64 CUDA tensor of the same logical shape.  An
    F-ordered 2-D numpy array (what sar_focus_csa returns) is uploaded through its C-ordered transpose and turned on the
    device instead of being re-strided on the host."""
    if torch.is_tensor(a):
        x = a if a.is_cuda else a.to(device)
        return x.contiguous() if x.dtype == torch.complex64 else dev.narrow_c128(x.to(torch.complex128).contiguous())
    h = np.asarray(a)
    turn = h.ndim == 2 and h.flags.f_contiguous and not h.flags.c_contiguous
    if turn:
        h = h.T
    x = hostio.to_device_c64(h, device)
    return dev.transpose_c32(x) if turn else x


# ------------------------------------------------------------------------------------------ echo
def run_bistatic_physics_gpu(targets, t_vec, pos_tx_np, vel_tx_np, rx_offset_dist, vel_target_np, *,
                             params: RadarParams | None = None, device=None, return_device=False):
    """Two-phase-centre echo (sar_ati_dcpa_sim_csa.py:106-181).  Returns (raw[P,S] complex128, t_start_fast)."""
    prm = _params(params)
    pos0, rcs = targets_to_arrays(targets)
    vel_tx = np.asarray(vel_tx_np, dtype=np.float64).reshape(-1, 3)
    pos_tx = np.asarray(pos_tx_np, dtype=np.float64).reshape(-1, 3)
    v_dir = vel_tx / np.sqrt(np.sum(vel_tx * vel_tx, axis=1))[:, None]           # :145
    pos_rx = pos_tx + v_dir * rx_offset_dist                                      # :148
    S = int(22e-6 * prm.FS) if prm.n_samples == 0 else prm.n_samples              # :111
    t0 = prm.t_start_fast                                                         # :112
    raw = dev.echo_accumulate(pos0, np.asarray(vel_target_np, dtype=np.float64).reshape(3), rcs, pos_tx, pos_rx,
                              np.asarray(t_vec, dtype=np.float64), c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p,
                              t_start=t0, fs=prm.FS, n_samples=S, device=device or _default_device)
    return (raw if return_device else _to_host_c128(raw)), t0


def run_physics_engine(targets, pos_sat, t_vec, *, params: RadarParams | None = None, device=None,
                       return_device=False):
    """Monostatic, static scatterers (sar_satellite_sim.py:211-305).  Returns (raw, t_start_fast, fs);
    fs is fixed at 600 MHz and S at int(22e-6 fs) inside the reference (:245-248)."""
    return run_moving_physics(targets, t_vec, pos_sat, (0.0, 0.0, 0.0), params=params, device=device,
                              return_device=return_device)


def run_moving_physics(targets, t_vec, pos_sat, vel_target, *, params: RadarParams | None = None, device=None,
                       return_device=False):
    """Monostatic, scatterers translating with ``vel_target`` (sar_satellite_moving_sim.py:111-159)."""
    prm = _params(params)
    fs = 600e6                                                                    # :114
    S = int(22e-6 * fs) if prm.n_samples == 0 else prm.n_samples
    t0 = (2 * prm.R0 / prm.C) - (prm.T_p / 2) - 1e-6                              # :116
    pos0, rcs = targets_to_arrays(targets)
    raw = dev.echo_accumulate(pos0, np.asarray(vel_target, dtype=np.float64).reshape(3), rcs,
                              np.asarray(pos_sat, dtype=np.float64), None, np.asarray(t_vec, dtype=np.float64),
                              c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, t_start=t0, fs=fs, n_samples=S,
                              device=device or _default_device)
    return (raw if return_device else _to_host_c128(raw)), t0, fs


def run_custom_physics(targets, t_vec, pos, tuned_prp, t_p, fc, bw, *, params: RadarParams | None = None,
                       device=None, return_device=False):
    """Airborne engine (sar_vehicle_sim.py:83-126): fs = 360 MHz, S = 2048, window centred on 2 R0/C.
    ``tuned_prp`` is unused by the reference as well.  Returns raw only."""
    prm = _params(params)
    fs, S = 360e6, 2048                                                           # :85-86
    t0 = (2 * prm.R0 / prm.C) - (S / fs) / 2                                      # :89
    pos0, rcs = targets_to_arrays(targets)
    raw = dev.echo_accumulate(pos0, np.zeros(3), rcs, np.asarray(pos, dtype=np.float64), None,
                              np.asarray(t_vec, dtype=np.float64), c=prm.C, fc=fc, k_rate=bw / t_p, t_p=t_p,
                              t_start=t0, fs=fs, n_samples=S, device=device or _default_device)
    return raw if return_device else _to_host_c128(raw)


# --------------------------------------------------------------------- spotlight echo + backprojection
def spotlight_window(prm: RadarParams):
    """Receive window of run_physics_spotlight (sar_batch_sim.py:85-90): (t_start, num_samples)."""
    win_len = (2000.0 / prm.C) + prm.T_p + 10e-6
    n = int(np.ceil(win_len * prm.FS))
    if n % 2 != 0:
        n += 1
    return 2 * prm.R0 / prm.C - win_len / 2, n


def run_physics_spotlight(base_targets, t_vec, pos_sat, vel_sat, heading_deg, speed, l_ant, *,
                          params: RadarParams | None = None, device=None, return_device=True):
    """Spotlight CPI echo of a target block on a heading (sar_batch_sim.py:83-169).  Returns
    (raw_sig, t_start, num_samples, v_tgt) -- raw_sig a complex64 CUDA tensor [P, S] (the reference returns a device
    tensor as well, complex128; ``return_device=False`` gives complex128 numpy)."""
    prm = _params(params)
    t_start, n = spotlight_window(prm)
    t_fast = t_start + np.arange(n) / prm.FS                                       # :90
    phi = np.radians(heading_deg)
    v_tgt = np.array([speed * np.cos(phi), speed * np.sin(phi), 0])                # :93
    c, s = np.cos(phi), np.sin(phi)
    rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    pos0 = np.array([rot @ np.asarray(t["position"], dtype=np.float64) for t in base_targets])   # :99
    rcs = np.array([t["rcs"] for t in base_targets], dtype=np.float64)
    raw = dev.echo_accumulate(pos0, v_tgt, rcs, np.asarray(pos_sat, dtype=np.float64), None,
                              np.asarray(t_vec, dtype=np.float64), c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p,
                              t_start=t_start, fs=prm.FS, n_samples=n, device=device or _default_device,
                              spotlight=(np.asarray(vel_sat, dtype=np.float64), np.pi * l_ant / prm.Lambda, t_fast))
    return (raw if return_device else _to_host_c128(raw)), t_start, n, v_tgt


_tdbp_plans: dict = {}


def tdbp_gpu(raw_t, pos_plat, vel_plat, t_start, num_samples, vel_focus, t_pulses, scene_size, nx=512, ny=512, *,
             params: RadarParams | None = None, device=None, return_device=False):
    """Time-domain backprojection of one CPI (sar_batch_sim.py:171-238): complex128 image [ny, nx]."""
    prm = _params(params)
    device = device or _default_device
    if torch.is_tensor(raw_t):
        x = raw_t if raw_t.dtype == torch.complex64 else dev.narrow_c128(raw_t.to(torch.complex128).contiguous())
        x = x.to(device) if not x.is_cuda else x
    else:
        x = hostio.to_device_c64(raw_t, device)
    key = (prm.C, prm.FC, prm.k_rate, prm.T_p, prm.FS, float(t_start), int(num_samples), float(scene_size), int(nx), int(ny),
           x.device.index)
    plan = _tdbp_plans.get(key)
    if plan is None:
        if len(_tdbp_plans) >= 4:
            _tdbp_plans.pop(next(iter(_tdbp_plans))).close()
        plan = dev.TdbpPlan(c=prm.C, fc=prm.FC, k_rate=prm.k_rate, t_p=prm.T_p, fs=prm.FS, t_start=float(t_start),
                            n_samples=int(num_samples), scene_size=float(scene_size), nx=nx, ny=ny, device=x.device)
        _tdbp_plans[key] = plan
    img = plan.backproject(plan.range_compress(x), pos_plat, vel_plat, t_pulses, vel_focus)
    return img if return_device else img.cpu().numpy()


# ------------------------------------------------------------------------------------------- CSA
def sar_focus_csa(phist, center_wavelength_m, pulse_width_sec, chirp_rate_hzpsec, sample_rate_hz, prf_hz,
                  platform_speed_mps, range_ref_m, t_start_fast, *, device=None, return_device=False, order="C",
                  out=None):
    """Chirp Scaling focusing (sar_ati_dcpa_sim_csa.py:202-396).  ``phist`` is [N_az, N_rg]: a numpy
    complex array (any complex dtype) or a complex64 CUDA tensor.  Returns (img, range_axis,
    cross_range_axis) with img of shape [N_rg, N_az] -- the array the reference returns as ``img.T`` --
    as complex128 numpy.  ``order="C"`` (default here): C-contiguous, the layout the device holds and every later
    stage wants; ``order="F"`` (what ``install()`` selects): the F-contiguous view semantics of the reference's
    ``img.T`` (:396), same values, same ``.flags`` / ``.strides``.  ``out``: optional complex128 array of that shape and
    order to receive the image (page-locked memory from ``nis_sar.hostio.pinned_empty`` avoids the staging copy).
    ``pulse_width_sec`` is unused, as in the reference."""
    if order not in ("C", "F"):
        raise dev.NisError("sar_focus_csa: order must be 'C' or 'F'")
    device = device or _default_device
    if torch.is_tensor(phist):
        x = phist if phist.dtype == torch.complex64 else dev.narrow_c128(phist.to(torch.complex128).contiguous())
    else:
        h = np.asarray(phist)
        if h.ndim != 2:
            raise dev.NisError("sar_focus_csa: phist must be 2-D [N_az, N_rg]")
        x = hostio.to_device_c64(h, device)
    n_az, n_rg = x.shape
    plan = dev.cached_plan(n_az, n_rg, lam=float(center_wavelength_m), kr=float(chirp_rate_hzpsec),
                           fs=float(sample_rate_hz), prf=float(prf_hz), vr=float(platform_speed_mps),
                           r_ref=float(range_ref_m), t_start=float(t_start_fast), device=x.device)
    slc = plan.focus(x)
    rax, cax = plan.axes()
    if return_device:
        return slc, rax, cax
    return (_to_host_c128_fortran(slc, out) if order == "F" else _to_host_c128(slc, out)), rax, cax


# ------------------------------------------------------------------------------------- noise / SNR
K_BOLTZ = 1.380649e-23
SNR_PRESETS = {
    "satellite": dict(p_tx=1000.0, ant_l=3.5, ant_w=0.5, t_sys=290.0, nf_db=5.0, loss_db=3.0),   # sar_satellite_sim.py:307-313
    "vehicle": dict(p_tx=2000.0, ant_l=1.5, ant_w=0.3, t_sys=290.0, nf_db=4.0, loss_db=3.0),     # sar_vehicle_sim.py:129-134
}


def calculate_snr_db(r_slant, rcs, wavelength, bandwidth, t_int, p_tx=None, ant_l=None, ant_w=None, t_sys=None,
                     nf_db=None, loss_db=None, *, preset="satellite"):
    """Radar-equation SNR (sar_satellite_sim.py:319-329; sar_vehicle_sim.py:140-150 with the airborne constants:
    ``preset="vehicle"``).  Returns (snr_db, gain_db) -- host scalar arithmetic, the same expressions as the reference."""
    d = dict(SNR_PRESETS[preset])
    for k, v in (("p_tx", p_tx), ("ant_l", ant_l), ("ant_w", ant_w), ("t_sys", t_sys), ("nf_db", nf_db), ("loss_db", loss_db)):
        if v is not None:
            d[k] = v
    ant_area = d["ant_l"] * d["ant_w"] * 0.6
    gain = 4 * np.pi * ant_area / (wavelength ** 2)
    gain_db = 10 * np.log10(gain)
    nf = 10 ** (d["nf_db"] / 10)
    loss = 10 ** (d["loss_db"] / 10)
    numerator = d["p_tx"] * (gain ** 2) * (wavelength ** 2) * rcs * t_int
    denominator = ((4 * np.pi) ** 3) * (r_slant ** 4) * K_BOLTZ * d["t_sys"] * bandwidth * loss * nf
    return 10 * np.log10(numerator / denominator), gain_db


def calculate_raw_snr_db(r_slant, rcs, wavelength, bandwidth, ant_l, p_tx=1000.0, ant_w=0.5, t_sys=290.0, nf_db=5.0,
                         loss_db=3.0):
    """Single-pulse (un-integrated) radar-equation SNR of sar_batch_sim.py:53-63 (its module constants as defaults)."""
    ant_area = ant_l * ant_w
    effective_area = ant_area * 0.6
    gain = 4 * np.pi * effective_area / (wavelength ** 2)
    nf = 10 ** (nf_db / 10)
    loss = 10 ** (loss_db / 10)
    numerator = p_tx * (gain ** 2) * (wavelength ** 2) * rcs
    denominator = ((4 * np.pi) ** 3) * (r_slant ** 4) * K_BOLTZ * t_sys * bandwidth * loss * nf
    return 10 * np.log10(numerator / denominator)


def _draw_seed(seed):
    # the reference draws from numpy's global generator; so does the default seed, which keeps np.random.seed() meaningful
    return int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + 12345 if seed is None else int(seed)


def add_ocean_noise(raw_data, snr_db, scr_db=10.0, k_nu=1.0, *, seed=None, device=None, return_device=False):
    """Thermal noise + K-distributed sea clutter (sar_satellite_sim.py:331-344).  A complex64 CUDA tensor is modified in
    place and returned (the echo stays in HBM between synthesis and focusing); a numpy array comes back as a new
    complex128 array, as in the reference.  Same distributions and powers as the reference; the draws themselves come from
    a counter-based generator keyed by ``seed`` (default: one draw from numpy's global generator)."""
    device = device or _default_device
    seed = _draw_seed(seed)
    if torch.is_tensor(raw_data):
        x = raw_data if raw_data.dtype == torch.complex64 else dev.narrow_c128(raw_data.to(torch.complex128).contiguous())
        return dev.add_noise(x.contiguous(), snr_db, scr_db, k_nu, seed)
    x = hostio.to_device_c64(raw_data, device)
    dev.add_noise(x, snr_db, scr_db, k_nu, seed)
    return x if return_device else _to_host_c128(x)


def generate_noise_tensor(shape, ref_power, snr_db, scr_db=10.0, k_nu=1.0, *, seed=None, device=None):
    """Noise + clutter alone, for a given reference power (sar_batch_sim.py:65-81): a complex64 CUDA tensor.
    ``ref_power``: a number or a 1-element CUDA tensor (e.g. ``nis_sar.device.peak_power(raw_sig)``, :317)."""
    device = device or (ref_power.device if torch.is_tensor(ref_power) and ref_power.is_cuda else _default_device)
    x = torch.empty(tuple(shape), dtype=torch.complex64, device=device)
    return dev.add_noise(x, snr_db, scr_db, k_nu, _draw_seed(seed),
                         ref_power=ref_power if torch.is_tensor(ref_power) else float(ref_power), accumulate=False)


# ------------------------------------------------------------------------------------------- RDA
_RDA_RETURNS = {
    # which of the three copies of sar_focus_rda the caller is replacing -> what it returns after the image and two axes
    "satellite": ("phist_compressed", "range_doppler", "range_doppler_rcmc"),                              # sar_satellite_sim.py:447-448
    "vehicle": ("phist_compressed", "range_doppler", "range_doppler_rcmc", "range_doppler_filtered"),    # sar_vehicle_sim.py:273-274
    "moving": (),                                                                                          # sar_satellite_moving_sim.py:285
}


def sar_focus_rda(phist, center_wavelength_m, pulse_width_sec, chirp_rate_hzpsec, sample_rate_hz, prf_hz,
                  platform_speed_mps, range_grp_m, *, returns="satellite", device=None, return_device=False):
    """Range-Doppler focusing (sar_satellite_sim.py:356-448 and its copies in sar_vehicle_sim.py:182-274,
    sar_satellite_moving_sim.py:208-285).  ``phist`` is [num_ranges, num_pulses] as in the reference -- normally the
    ``raw_data.T`` view of the pulse-major echo array, which is consumed without a copy -- or a complex64 CUDA tensor of
    that shape.  ``returns`` selects which copy's tuple comes back:
      "satellite": (image_mag_T, range_axis_centered, cross_range, phist_compressed, range_doppler, range_doppler_rcmc,
                    doppler_freq);  "vehicle": the same plus range_doppler_filtered before doppler_freq;  "moving": the
      first three.  image_mag_T is float64 [num_pulses, num_ranges]; the complex arrays are complex128
      [num_ranges, num_pulses] (transposed views of pulse-major buffers: same values, F-contiguous)."""
    if returns not in _RDA_RETURNS:
        raise dev.NisError(f"sar_focus_rda: returns must be one of {sorted(_RDA_RETURNS)}")
    device = device or _default_device
    if torch.is_tensor(phist):
        xt = phist.transpose(0, 1)
        if xt.dtype != torch.complex64:
            xt = dev.narrow_c128(xt.to(torch.complex128).contiguous())
        elif xt.stride(1) != 1:
            xt = xt.contiguous()
    else:
        h = np.asarray(phist)
        if h.ndim != 2:
            raise dev.NisError("sar_focus_rda: phist must be 2-D [num_ranges, num_pulses]")
        ht = h.T                                     # pulse-major; C-contiguous when phist is raw_data.T
        xt = hostio.to_device_c64(ht, device)
    n_pulses, n_ranges = xt.shape
    plan = dev.cached_rda_plan(n_pulses, n_ranges, lam=float(center_wavelength_m), t_p=float(pulse_width_sec),
                               kr=float(chirp_rate_hzpsec), fs=float(sample_rate_hz), prf=float(prf_hz),
                               vr=float(platform_speed_mps), range_grp=float(range_grp_m), device=xt.device)
    want = _RDA_RETURNS[returns]
    out = plan.focus(xt, want=want)
    rax, cax, dop = plan.axes()
    if return_device:
        img = out["image_mag"]
        extra = [out[w].transpose(0, 1) for w in want]
    else:
        img = out["image_mag"].to(torch.float64).cpu().numpy()
        extra = [_to_host_c128(out[w]).T for w in want]
    if returns == "moving":
        return img, rax, cax
    return (img, rax, cax, *extra, dop)


# ------------------------------------------------------------------------------------------ GMTI
def gmti_products(slc1, slc2, thresh=0.05, cal_phase=0.0, *, device=None, return_device=False):
    """ATI interferogram, phase, DPCA difference, magnitudes, 5 %-of-peak mask, masked phase and the
    detected-pixel list (sar_ati_dcpa_sim_csa.py:414-419, :447-449).  Inputs [N_rg, N_az] complex."""
    device = device or _default_device

    out = dev.gmti_fused(_upload_c32(slc1, device), _upload_c32(slc2, device), thresh, cal_phase)
    if return_device:
        return out
    res = {}
    for k, v in out.items():
        if not torch.is_tensor(v):
            res[k] = v
        elif v.dtype == torch.complex64:
            res[k] = _to_host_c128(v)
        elif v.dtype == torch.float32:
            res[k] = v.to(torch.float64).cpu().numpy()
        else:
            res[k] = v.cpu().numpy()
    return res


def dpca_coregister(raw_rx1, raw_rx2):
    """rx1[1:], rx2[:-1] (sar_ati_dcpa_sim_csa.py:402-403): views, no copy (numpy or torch)."""
    return raw_rx1[1:, :], raw_rx2[:-1, :]


def save_ati_dpca_npz(fname, slc1, slc2, range_axis, cross_range):
    """The hand-off file of the reference (sar_ati_dcpa_sim_csa.py:457-461): keys slc1, slc2 ([N_rg, N_az]
    complex128), range_axis, cross_range -- what sar_ati_dcpa_viewer_csa.py:24-29 loads and transposes back."""
    np.savez(fname, slc1=np.asarray(slc1, dtype=np.complex128), slc2=np.asarray(slc2, dtype=np.complex128),
             range_axis=np.asarray(range_axis, dtype=np.float64), cross_range=np.asarray(cross_range, dtype=np.float64))


# --------------------------------------------------------------------------------------- install
_ENTRY_POINTS = ("run_bistatic_physics_gpu", "sar_focus_csa", "run_physics_engine", "run_moving_physics",
                 "run_custom_physics")
_INSTALLABLE = _ENTRY_POINTS + ("sar_focus_rda", "add_ocean_noise", "calculate_snr_db", "run_physics_spotlight", "tdbp_gpu",
                                "generate_noise_tensor", "calculate_raw_snr_db")


def install(namespace: dict, names=_ENTRY_POINTS, *, rda_returns="satellite", snr_preset="satellite"):
    """Patch the reference's entry points inside ``namespace`` (a simulator module's ``globals()``) with the CUDA
    implementations -- the drop-in mechanism INTEGRATION.md describes.  Replacements of functions that read radar
    constants from their module (``C, R0, FC, BW, T_p, FS``: sar_ati_dcpa_sim_csa.py:111-115, :159-168) read them from
    ``namespace`` on every call, exactly like the functions they replace; functions that take everything as arguments
    are installed as they are.  ``sar_focus_csa`` is installed with ``order="F"`` (the reference's ``img.T`` view
    semantics, :396).  ``rda_returns`` picks which copy of ``sar_focus_rda`` is being replaced ("satellite": 7-tuple,
    "vehicle": 8-tuple, "moving": 3-tuple); ``snr_preset`` the radar constants of ``calculate_snr_db``."""
    import functools
    import inspect

    def live_params():
        base = spaceborne_preset()
        kw = {k: float(namespace[k]) for k in ("C", "R0", "FC", "BW", "T_p") if k in namespace}
        if "FS" in namespace:
            kw["FS"] = float(namespace["FS"])
        return base.replace(**kw)

    def with_live_params(fn):
        @functools.wraps(fn)
        def patched(*a, **k):
            k.setdefault("params", live_params())
            return fn(*a, **k)
        return patched

    g = globals()
    fixed = {"sar_focus_csa": {"order": "F"}, "sar_focus_rda": {"returns": rda_returns},
             "calculate_snr_db": {"preset": snr_preset}}
    for n in names:
        if n not in _INSTALLABLE:
            raise dev.NisError(f"install: {n!r} is not a replaceable entry point (known: {', '.join(_INSTALLABLE)})")
        fn = g[n]
        if n in fixed:
            fn = functools.wraps(fn)(functools.partial(fn, **fixed[n]))
        if "params" in inspect.signature(g[n]).parameters:
            fn = with_live_params(fn)
        namespace[n] = fn
    return namespace
