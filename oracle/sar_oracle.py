"""TEST INFRASTRUCTURE ONLY -- numpy (fp64) restatement of the reference hot path.

Each function names the reference lines whose arithmetic it follows.  It is a
restatement (written from the maths, broadcasting instead of meshgrid copies),
not a copy; ``tests/test_oracle_golden.py`` pins it to vectors produced by the
reference's own functions (``oracle/make_golden.py``).

Radar constants arrive as a plain dict ``g`` with the names of the reference's
module globals: ``C, R0, FC, BW, T_p, FS`` (sar_ati_dcpa_sim_csa.py:18-38).
"""
from __future__ import annotations

import numpy as np

C_LIGHT = 299792458.0


# --------------------------------------------------------------------------- echo
def fast_time_axis(t_start: float, n_samples: int, fs: float) -> np.ndarray:
    """Receive-window sample times.  The grid is ``linspace(0, S/fs, S)`` -- S points
    *including* the end point, so the step is S/((S-1) fs), not 1/fs
    (sar_ati_dcpa_sim_csa.py:113-114; sar_satellite_sim.py:254-255;
    sar_vehicle_sim.py:88,100)."""
    return t_start + np.linspace(0.0, n_samples / fs, n_samples)


def window_start_satellite(g: dict) -> float:
    """Window opens 1 us + T_p/2 before the scene-centre echo
    (sar_ati_dcpa_sim_csa.py:112; sar_satellite_sim.py:252)."""
    return (2 * g["R0"] / g["C"]) - (g["T_p"] / 2) - 1e-6


def _accumulate_pulse(t_fast, tau, phase0, amp, t_p, k_rate, block=256):
    """One pulse row: sum_b amp_b exp(j(phase0_b + pi k (t - tau_b - T_p/2)^2)) gated to the
    closed interval |t - tau_b - T_p/2| <= T_p/2 (sar_ati_dcpa_sim_csa.py:163-176)."""
    row = np.zeros(t_fast.shape[0], dtype=np.complex128)
    half = t_p / 2
    for b0 in range(0, tau.shape[0], block):
        sl = slice(b0, b0 + block)
        u = (t_fast[None, :] - tau[sl, None]) - half
        gate = np.abs(u) <= half
        arg = phase0[sl, None] + np.pi * k_rate * (u ** 2)
        row += np.sum(amp[sl, None] * np.exp(1j * arg) * gate, axis=0)
    return row


def echo_bistatic(pos0, rcs, t_slow, pos_tx, vel_tx, rx_offset, vel_target, g, n_samples=None):
    """Two-phase-centre chirp echo of ``run_bistatic_physics_gpu``
    (sar_ati_dcpa_sim_csa.py:106-181).  Receiver sits ``rx_offset`` metres along the
    unit velocity from the transmitter (:145-148); every scatterer of the call moves
    with the one ``vel_target`` (:151); delay is (|p-p_tx| + |p-p_rx|)/C (:154-159);
    carrier phase -2 pi FC tau (:160); amplitude sqrt(rcs) (:171).
    Returns (raw[P, S] complex128, t_start_fast)."""
    pos0 = np.asarray(pos0, dtype=np.float64).reshape(-1, 3)
    rcs = np.asarray(rcs, dtype=np.float64).reshape(-1)
    vel_target = np.asarray(vel_target, dtype=np.float64).reshape(3)
    S = int(22e-6 * g["FS"]) if n_samples is None else int(n_samples)
    t0 = window_start_satellite(g)
    t_fast = fast_time_axis(t0, S, g["FS"])
    k_rate = g["BW"] / g["T_p"]
    amp = np.sqrt(rcs)
    raw = np.zeros((len(t_slow), S), dtype=np.complex128)
    for i in range(len(t_slow)):
        p_tx = np.asarray(pos_tx[i], dtype=np.float64)
        v = np.asarray(vel_tx[i], dtype=np.float64)
        p_rx = p_tx + (v / np.sqrt(np.sum(v * v))) * rx_offset
        p = pos0 + vel_target[None, :] * t_slow[i]
        d_tx = np.sqrt(np.sum((p - p_tx) ** 2, axis=1))
        d_rx = np.sqrt(np.sum((p - p_rx) ** 2, axis=1))
        tau = (d_tx + d_rx) / g["C"]
        raw[i] = _accumulate_pulse(t_fast, tau, -2.0 * np.pi * g["FC"] * tau, amp, g["T_p"], k_rate)
    return raw, t0


def echo_monostatic(pos0, rcs, t_slow, pos_sat, g, vel_target=None, n_samples=None,
                    fs=None, t_start=None, t_p=None, fc=None, bw=None):
    """Monostatic engines: ``run_physics_engine`` (sar_satellite_sim.py:211-305, static
    scatterers), ``run_moving_physics`` (sar_satellite_moving_sim.py:111-159, p = p0 + v t
    at :138) and ``run_custom_physics`` (sar_vehicle_sim.py:83-126, fs=360 MHz, S=2048,
    window centred on 2 R0/C at :89).  tau = 2 d / C, carrier phase -4 pi FC d / C
    (sar_satellite_sim.py:273-274)."""
    pos0 = np.asarray(pos0, dtype=np.float64).reshape(-1, 3)
    rcs = np.asarray(rcs, dtype=np.float64).reshape(-1)
    fs = 600e6 if fs is None else fs
    t_p = g["T_p"] if t_p is None else t_p
    fc = g["FC"] if fc is None else fc
    bw = g["BW"] if bw is None else bw
    S = int(22e-6 * fs) if n_samples is None else int(n_samples)
    t0 = window_start_satellite(g) if t_start is None else t_start
    t_fast = fast_time_axis(t0, S, fs)
    k_rate = bw / t_p
    amp = np.sqrt(rcs)
    vt = None if vel_target is None else np.asarray(vel_target, dtype=np.float64).reshape(3)
    raw = np.zeros((len(t_slow), S), dtype=np.complex128)
    for i in range(len(t_slow)):
        p = pos0 if vt is None else pos0 + vt * t_slow[i]
        d = np.sqrt(np.sum((p - np.asarray(pos_sat[i], dtype=np.float64)) ** 2, axis=1))
        tau = 2 * d / g["C"]
        raw[i] = _accumulate_pulse(t_fast, tau, -4.0 * np.pi * fc * d / g["C"], amp, t_p, k_rate)
    return raw, t0, fs


def vehicle_window_start(g: dict, n_samples: int = 2048, fs: float = 360e6) -> float:
    """sar_vehicle_sim.py:89 -- window centred on the scene-centre delay."""
    return (2 * g["R0"] / g["C"]) - (n_samples / fs) / 2


# ---------------------------------------------------------------------------- CSA
def csa_axes(n_az, n_rg, lam, fs, prf, vr, r_ref, t_start):
    """Axes and per-Doppler scalars of ``sar_focus_csa`` (sar_ati_dcpa_sim_csa.py:217-262):
    tau_n = t_start + n/fs (:219), shifted frequency axes (:222-235, :281), migration factor
    D(fa) = sqrt(1 - (lam fa / 2 Vr)^2) with negative arguments clamped to 1e-9 (:244-247),
    Cs = 1/D - 1 (:249), tau_ref = 2 R_ref / (c D) (:262)."""
    dt = 1.0 / fs
    tau = t_start + np.arange(n_rg) * dt
    fr = np.fft.fftshift(np.fft.fftfreq(n_rg, dt))
    fa = np.fft.fftshift(np.fft.fftfreq(n_az, 1.0 / prf))
    arg = 1.0 - (lam * fa / (2.0 * vr)) ** 2
    arg[arg < 0] = 1e-9
    d = np.sqrt(arg)
    cs = (1.0 / d) - 1.0
    tau_ref = 2.0 * r_ref / (C_LIGHT * d)
    return tau, fr, fa, d, cs, tau_ref


def focus_csa(phist, lam, kr, fs, prf, vr, r_ref, t_start):
    """Chirp Scaling focusing, ``sar_focus_csa`` (sar_ati_dcpa_sim_csa.py:202-396).

    azimuth FFT + shift (:233-234) -> x Phi1 = exp(-j pi Kr Cs (tau - tau_ref)^2) (:272-274)
    -> range FFT + shift (:278-280) -> x Phi2 = exp(j(pi fr^2/(Kr(1+Cs)) + 4 pi R_ref Cs fr/c))
    (:318-326) -> unshift + range IFFT (:331) -> x Phi3 = exp(j(4 pi R D/lam
    - pi Kr Cs(1+Cs)(tau - 2 R_ref/c)^2)), R = c tau/2 (:346-382) -> unshift + azimuth IFFT (:385).
    Returns (img.T  -- an [N_rg, N_az] view --, range_axis, cross_range_axis) (:392-396)."""
    phist = np.asarray(phist)
    n_az, n_rg = phist.shape
    tau, fr, fa, d, cs, tau_ref = csa_axes(n_az, n_rg, lam, fs, prf, vr, r_ref, t_start)
    c = C_LIGHT
    csc = cs[:, None]

    s = np.fft.fftshift(np.fft.fft(phist, axis=0), axes=0)
    s = s * np.exp(-1j * np.pi * kr * csc * (tau[None, :] - tau_ref[:, None]) ** 2)

    s = np.fft.fftshift(np.fft.fft(s, axis=1), axes=1)
    frr = fr[None, :]
    s = s * np.exp(1j * (np.pi * (frr ** 2) / (kr * (1.0 + csc)) + 4.0 * np.pi * r_ref * csc * frr / c))

    s = np.fft.ifft(np.fft.ifftshift(s, axes=1), axis=1)
    rng = c * tau / 2.0
    resid = -np.pi * kr * csc * (1.0 + csc) * ((tau[None, :] - (2.0 * r_ref / c)) ** 2)
    s = s * np.exp(1j * (4.0 * np.pi * rng[None, :] * d[:, None] / lam + resid))

    img = np.fft.ifft(np.fft.ifftshift(s, axes=0), axis=0)
    t_slow = np.arange(n_az) / prf
    t_slow -= np.mean(t_slow)
    return img.T, rng, t_slow * vr


# --------------------------------------------------------------------------- GMTI
def dpca_coregister(raw_rx1, raw_rx2):
    """One-pulse shift that aligns the two phase centres (sar_ati_dcpa_sim_csa.py:402-403)."""
    return raw_rx1[1:, :], raw_rx2[:-1, :]


def gmti_products(slc1, slc2, thresh_frac=0.05, cal_phase=0.0):
    """ATI / DPCA products (sar_ati_dcpa_sim_csa.py:414-419, :447-449) with the viewer's
    optional channel balance s2 * exp(j cal) (sar_ati_dcpa_viewer_csa.py:43).
    Detected pixels = flatnonzero(|slc1| > thresh_frac * max|slc1|) in row-major order of
    the [N_rg, N_az] array; peak = first argmax |slc1|."""
    slc1 = np.asarray(slc1)
    s2 = np.asarray(slc2)
    if cal_phase != 0.0:
        s2 = s2 * np.exp(1j * cal_phase)
    interf = slc1 * np.conj(s2)
    phase = np.angle(interf)
    mag1 = np.abs(slc1)
    diff = slc1 - s2
    dmag = np.abs(diff)
    mask = mag1 > (np.max(mag1) * thresh_frac)
    phase_masked = np.copy(phase)
    phase_masked[~mask] = 0
    return {
        "ati_interf": interf, "ati_phase": phase, "slc1_mag": mag1,
        "dpca_diff": diff, "dpca_mag": dmag, "mag_mask": mask,
        "ati_phase_masked": phase_masked,
        "det_idx": np.flatnonzero(mask),
        "peak_idx": int(np.argmax(mag1)),
    }


def balance_phase(slc1, slc2):
    """Viewer auto-balance: angle(mean(slc1 conj(slc2))) (sar_ati_dcpa_viewer_csa.py:249-250)."""
    return float(np.angle(np.mean(np.asarray(slc1) * np.conj(np.asarray(slc2)))))


def threshold_margin(slc1, thresh_frac=0.05):
    """Smallest relative distance of any pixel magnitude to the detection threshold --
    reported next to every bit-exact index comparison so a flip is attributable."""
    mag = np.abs(np.asarray(slc1)).ravel()
    thr = mag.max() * thresh_frac
    return float(np.min(np.abs(mag - thr)) / thr)


# ---------------------------------------------------------------------------- RDA
def hamming_sym(m: int) -> np.ndarray:
    """scipy.signal.windows.hamming(m) (symmetric): 0.54 - 0.46 cos(2 pi n / (m - 1))."""
    if m == 1:
        return np.ones(1)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(m) / (m - 1))


def rda_axes(num_ranges, num_pulses, fs, prf, range_grp_m):
    """Axes of ``sar_focus_rda`` (sar_satellite_sim.py:363-375, :400-406): slow time and fast time centred
    on sample n/2 (even) or (n-1)/2 (odd), Doppler axis k PRF / P in fftshift order, range = c t / 2."""
    c = 299792458
    if num_pulses % 2 == 0:
        slow = (np.arange(num_pulses) - num_pulses / 2) / prf
        dop = np.arange(-num_pulses / 2, num_pulses / 2) * (prf / num_pulses)
    else:
        slow = (np.arange(num_pulses) - (num_pulses - 1) / 2) / prf
        dop = np.arange(-(num_pulses - 1) / 2, (num_pulses - 1) / 2 + 1) * (prf / num_pulses)
    t_grp = 2 * range_grp_m / c
    if num_ranges % 2 == 0:
        fast = (np.arange(num_ranges) - num_ranges / 2) / fs + t_grp
    else:
        fast = (np.arange(num_ranges) - (num_ranges - 1) / 2) / fs + t_grp
    return slow, fast, dop, fast * c / 2


def rda_matched_filter(t_p, kr, fs):
    """Hamming-weighted, unit-norm matched filter (sar_satellite_sim.py:379-386):
    L = floor(T_p / (1/fs)) + 1 taps on linspace(-T_p/2, T_p/2, L)."""
    step = 1 / fs
    n_mf = int(np.floor(t_p / step)) + 1
    t = np.linspace(-t_p / 2, t_p / 2, n_mf)
    mf = np.conj(np.exp(1j * np.pi * kr * t ** 2)) * hamming_sym(n_mf)
    return mf / np.linalg.norm(mf)


def convolve_same(x, h):
    """scipy.signal.convolve(x, h, mode='same') along axis 0 of x ([N, P]): the N samples of the full linear
    convolution starting at (len(h) - 1) // 2, here through zero-padded FFTs (exact to fp64 rounding)."""
    n, l = x.shape[0], len(h)
    m = 1
    while m < n + l - 1:
        m *= 2
    y = np.fft.ifft(np.fft.fft(x, m, axis=0) * np.fft.fft(h, m)[:, None], axis=0)
    s = (l - 1) // 2
    return y[s:s + n]


def interp_linear_zero(x, y, xn):
    """scipy.interpolate.interp1d(x, y, kind='linear', fill_value=0, bounds_error=False)(xn) for increasing x:
    neighbours from searchsorted (left), clipped to [1, n-1]; zero outside [x[0], x[-1]]."""
    idx = np.clip(np.searchsorted(x, xn), 1, len(x) - 1)
    lo, hi = idx - 1, idx
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    out = slope * (xn - x[lo]) + y[lo]
    out[(xn < x[0]) | (xn > x[-1])] = 0
    return out


def focus_rda(phist, lam, t_p, kr, fs, prf, vr, range_grp_m):
    """Range-Doppler focusing, ``sar_focus_rda`` (sar_satellite_sim.py:356-448; the copies at
    sar_vehicle_sim.py:182-274 and sar_satellite_moving_sim.py:208-285 differ only in what they return).
    phist is [num_ranges, num_pulses].  matched-filter range compression (:377-392) -> azimuth Hamming, shifted
    FFT (:396-398) -> RCMC: per Doppler column, linear interpolation from the axis r (1 - fd^2 lam^2 / 8 Vr^2)
    back onto r, zero outside (:408-427) -> x exp(-j pi fd^2 / Ka), Ka = 2 Vr^2 / (lam r) (:431-435) -> shifted
    azimuth IFFT (:438).  Returns a dict of every array any of the three variants returns."""
    phist = np.asarray(phist)
    num_ranges, num_pulses = phist.shape
    slow, fast, dop, rax = rda_axes(num_ranges, num_pulses, fs, prf, range_grp_m)
    rc = convolve_same(phist.astype(complex), rda_matched_filter(t_p, kr, fs))
    rd = np.fft.fftshift(np.fft.fft(np.fft.fftshift(rc * hamming_sym(num_pulses), axes=1), axis=1), axes=1)
    c = 299792458
    lambd = c / (c / lam)
    rcmc = np.zeros_like(rd)
    for k in range(num_pulses):
        d_r = rax * (dop[k] ** 2) * lambd ** 2 / (8 * vr ** 2)
        rcmc[:, k] = interp_linear_zero(rax - d_r, rd[:, k], rax) if num_ranges > 1 else rd[:, k]
    inv_ka = 1.0 / ((2 * vr ** 2) / (lambd * rax))
    filt = rcmc * np.exp(-1j * np.pi * (inv_ka[:, None] * dop[None, :] ** 2))
    img = np.fft.ifftshift(np.fft.ifft(np.fft.ifftshift(filt, axes=1), axis=1), axes=1)
    return {"image_mag_T": np.abs(img).T, "range_axis_centered": rax - np.mean(rax), "cross_range": vr * slow,
            "phist_compressed": rc, "range_doppler": rd, "range_doppler_rcmc": rcmc, "range_doppler_filtered": filt,
            "doppler_freq": dop}


# ---------------------------------------------------------------------------- noise / clutter
SNR_PRESETS = {
    # module constants the three copies of calculate_snr_db take as defaults
    "satellite": dict(p_tx=1000.0, ant_l=3.5, ant_w=0.5, t_sys=290.0, nf_db=5.0, loss_db=3.0),   # sar_satellite_sim.py:307-313
    "vehicle": dict(p_tx=2000.0, ant_l=1.5, ant_w=0.3, t_sys=290.0, nf_db=4.0, loss_db=3.0),     # sar_vehicle_sim.py:129-134
}
K_BOLTZ = 1.380649e-23


def calculate_snr_db(r_slant, rcs, wavelength, bandwidth, t_int, p_tx, ant_l, ant_w, t_sys, nf_db, loss_db):
    """Radar-equation SNR after integration over t_int (sar_satellite_sim.py:319-329).  Returns (snr_db, gain_db)."""
    ant_area = ant_l * ant_w * 0.6
    gain = 4 * np.pi * ant_area / (wavelength ** 2)
    gain_db = 10 * np.log10(gain)
    nf = 10 ** (nf_db / 10)
    loss = 10 ** (loss_db / 10)
    numerator = p_tx * (gain ** 2) * (wavelength ** 2) * rcs * t_int
    denominator = ((4 * np.pi) ** 3) * (r_slant ** 4) * K_BOLTZ * t_sys * bandwidth * loss * nf
    return 10 * np.log10(numerator / denominator), gain_db


def ocean_noise(raw_data, snr_db, scr_db=10.0, k_nu=1.0, rng=None):
    """``add_ocean_noise`` (sar_satellite_sim.py:331-344): complex Gaussian thermal noise at signal_power / SNR plus
    K-distributed sea clutter (Gamma(nu, 1/nu) texture x Exp(1) speckle, uniform phase) at signal_power / SCR.
    ``rng``: a numpy RandomState drawn from in the reference's order (randn, randn, gamma, exponential, uniform);
    RandomState(s) reproduces the reference after np.random.seed(s) bit for bit."""
    rng = rng or np.random.RandomState()
    signal_power = np.mean(np.abs(raw_data) ** 2)
    noise_power = signal_power / (10 ** (snr_db / 10))
    thermal = np.sqrt(noise_power / 2) * (rng.randn(*raw_data.shape) + 1j * rng.randn(*raw_data.shape))
    clutter_power = signal_power / (10 ** (scr_db / 10))
    texture = rng.gamma(k_nu, 1 / k_nu, raw_data.shape)
    speckle = rng.exponential(1, raw_data.shape)
    phase = rng.uniform(0, 2 * np.pi, raw_data.shape)
    clutter = np.sqrt(clutter_power * texture * speckle) * np.exp(1j * phase)
    return raw_data + thermal + clutter, {"signal_power": signal_power, "noise_power": noise_power,
                                          "clutter_power": clutter_power}


# ---------------------------------------------------------------------------- viewer data layer
VIEWER_MODES = ("Ch1 Magnitude", "Ch1 Phase", "Ch2 Magnitude", "Ch2 Phase", "DPCA Magnitude", "DPCA Phase", "ATI Phase")


def viewer_products(s1, s2, cal_phase=0.0):
    """``SARData.compute_all`` (sar_ati_dcpa_viewer_csa.py:42-52)."""
    s2_cal = s2 * np.exp(1j * cal_phase)
    return {"Ch1 Magnitude": np.abs(s1), "Ch1 Phase": np.angle(s1), "Ch2 Magnitude": np.abs(s2_cal),
            "Ch2 Phase": np.angle(s2_cal), "DPCA Magnitude": np.abs(s1 - s2_cal), "DPCA Phase": np.angle(s1 - s2_cal),
            "ATI Phase": np.angle(s1 * np.conj(s2_cal))}


def viewer_visible_stats(prods, mode, scale, c_indices, r_indices):
    """The numbers ``print_visible_stats`` prints (sar_ati_dcpa_viewer_csa.py:103-145) and the colour limit it sets
    (:147-151) for the rectangle ``np.ix_(c_indices, r_indices)`` of the [N_cross, N_range] map."""
    vis = prods[mode][np.ix_(c_indices, r_indices)]
    if "Phase" in mode:
        data = vis
    else:
        data = 20 * np.log10(vis + 1e-12) if scale == "dB" else vis
    out = {"mean": np.mean(data), "median": np.median(data), "std": np.std(data), "min": np.min(data), "max": np.max(data)}
    if "DPCA" in mode:
        ref = prods["Ch1 Magnitude"][np.ix_(c_indices, r_indices)]
        out["cancellation_ratio"] = np.mean(ref) / (np.mean(vis) + 1e-9)
    if "Phase" in mode:
        out["clim"] = (-np.pi, np.pi)
    else:
        vmax = np.percentile(data, 99.9)
        out["clim"] = (vmax - 60 if scale == "dB" else 0, vmax)
    return out


# ---------------------------------------------------------------------------- spotlight echo + backprojection (8f, N4)
def spotlight_window(g):
    """Receive window of ``run_physics_spotlight`` (sar_batch_sim.py:85-90): 2 km of swath + pulse + 10 us, an even
    number of samples, centred on the scene-centre delay.  g: dict with C, R0, T_P, FS."""
    win_len = (2000.0 / g["C"]) + g["T_P"] + 10e-6
    n = int(np.ceil(win_len * g["FS"]))
    if n % 2 != 0:
        n += 1
    t_start = 2 * g["R0"] / g["C"] - win_len / 2
    return t_start, n


def echo_spotlight(pos0, rcs, t_vec, pos_sat, vel_sat, heading_deg, speed, l_ant, g):
    """``run_physics_spotlight`` (sar_batch_sim.py:83-169).  Target block rotated by the heading and moving at
    ``speed`` along it (:92-100); start-stop corrected two-way delay: the receive position is the platform advanced by
    v_sat * 2 d_tx / c (:133-137); one-way sinc^2 pattern of an aperture l_ant steered at the scene centre (:139-149);
    amplitude rcs * gain (NOT sqrt); chirp centred ON the delay: pi K (t - tau)^2 - 2 pi FC tau for |t - tau| <= T_P/2
    (:151-155).  g: dict with C, R0, FC, T_P, K_RATE, FS, Lambda.  Returns (raw[P, S], t_start, S, v_tgt)."""
    t_start, n = spotlight_window(g)
    t_fast = t_start + np.arange(n) / g["FS"]
    phi = np.radians(heading_deg)
    v_tgt = np.array([speed * np.cos(phi), speed * np.sin(phi), 0])
    c, s = np.cos(phi), np.sin(phi)
    rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    p0 = np.array([rot @ np.asarray(p, dtype=float) for p in pos0])
    rcs = np.asarray(rcs, dtype=float)
    raw = np.zeros((len(t_vec), n), dtype=complex)
    for i, t in enumerate(t_vec):
        p_tgt = p0 + v_tgt[None, :] * t
        diff_tx = p_tgt - pos_sat[i][None, :]
        dist_tx = np.linalg.norm(diff_tx, axis=1)
        p_rx = pos_sat[i][None, :] + vel_sat[i][None, :] * (2 * dist_tx / g["C"])[:, None]
        dist_rx = np.linalg.norm(p_tgt - p_rx, axis=1)
        tau = (dist_tx + dist_rx) / g["C"]
        b_vec = -pos_sat[i]
        look = b_vec / np.linalg.norm(b_vec)
        cos_off = np.clip((diff_tx / dist_tx[:, None]) @ look, -1, 1)
        x = np.pi * l_ant * np.sin(np.arccos(cos_off)) / g["Lambda"]
        gain = np.ones_like(x)
        m = np.abs(x) > 1e-6
        gain[m] = (np.sin(x[m]) / x[m]) ** 2
        t_local = t_fast[None, :] - tau[:, None]
        mask = np.abs(t_local) <= (g["T_P"] / 2)
        phase = np.pi * g["K_RATE"] * t_local ** 2 - 2 * np.pi * g["FC"] * tau[:, None]
        raw[i] = np.sum((rcs * gain)[:, None] * np.exp(1j * phase) * mask, axis=0)
    return raw, t_start, n, v_tgt


def tdbp_range_compress(raw, g, num_samples):
    """Circular matched filtering of ``tdbp_gpu`` (sar_batch_sim.py:179-185): the reference chirp has int(T_P FS) taps on
    linspace(-T_P/2, T_P/2), is fftshifted, zero-padded to num_samples, and correlated through length-num_samples FFTs."""
    n_ref = int(g["T_P"] * g["FS"])
    t_ref = np.linspace(-g["T_P"] / 2, g["T_P"] / 2, n_ref)
    ref_f = np.fft.fft(np.fft.fftshift(np.exp(1j * np.pi * g["K_RATE"] * t_ref ** 2)), n=num_samples)
    return np.fft.ifft(np.fft.fft(raw, n=num_samples, axis=1) * np.conj(ref_f)[None, :], axis=1)


def grid_sample_row_f32(row_f32, x_norm_f32):
    """torch.nn.functional.grid_sample(input[1, 1, 1, W], grid(x, y=0), bilinear, zeros padding, align_corners=False)
    as the CPU kernel evaluates it in float32: ix = fma(x + 1, W, -1) / 2 (ONE rounding of the multiply-subtract --
    pinned against torch in oracle/make_golden.py), neighbours floor(ix) and +1 with weights (i0 + 1 - ix), (ix - i0),
    out-of-range neighbours contribute zero."""
    f32 = np.float32
    w = row_f32.shape[-1]
    xp1 = (x_norm_f32.astype(f32) + f32(1)).astype(f32)
    ix = ((xp1.astype(np.float64) * w - 1.0).astype(f32) / f32(2)).astype(f32)    # the fp64 product is exact: one rounding
    i0 = np.floor(ix)
    w1 = (ix - i0).astype(f32)
    w0 = ((i0 + f32(1)) - ix).astype(f32)
    i0 = i0.astype(np.int64)
    i1 = i0 + 1
    a = np.where((i0 >= 0) & (i0 < w), row_f32[np.clip(i0, 0, w - 1)], f32(0)).astype(row_f32.dtype)
    b = np.where((i1 >= 0) & (i1 < w), row_f32[np.clip(i1, 0, w - 1)], f32(0)).astype(row_f32.dtype)
    return a * w0 + b * w1


def tdbp(raw, pos_plat, vel_plat, t_start, num_samples, vel_focus, t_pulses, scene_size, g, nx=512, ny=512):
    """Time-domain backprojection, ``tdbp_gpu`` (sar_batch_sim.py:171-238).  Per pixel and pulse (fp64): the pixel moves
    with vel_focus about the CPI centre (:207-209); two-way delay with the start-stop correction on both ends (:219-223);
    a range-Doppler coupling shift -FC (2 v_rad / C) / K_RATE (:214-217); the range-compressed pulse is sampled at
    (tau - t_start + t_shift) FS - 0.5 by float32 linear interpolation (grid_sample, align_corners=False, :225-230);
    x exp(j 2 pi FC tau), summed over pulses in complex128 (:232-235).  Returns [ny, nx]."""
    rc = tdbp_range_compress(np.asarray(raw), g, num_samples)
    rc_re, rc_im = rc.real.astype(np.float32), rc.imag.astype(np.float32)
    x_axis = np.linspace(-scene_size / 2, scene_size / 2, nx)
    y_axis = np.linspace(-scene_size / 2, scene_size / 2, ny)
    gx, gy = np.meshgrid(x_axis, y_axis, indexing="xy")
    grid = np.stack((gx.ravel(), gy.ravel(), np.zeros(gx.size)), axis=1)
    v_f = np.asarray(vel_focus, dtype=float)
    t_p = np.asarray(t_pulses, dtype=float)
    dt = t_p - np.mean(t_p)
    img = np.zeros(grid.shape[0], dtype=complex)
    for k in range(len(t_p)):
        gk = grid + v_f[None, :] * dt[k]
        diff_tx = gk - pos_plat[k][None, :]
        dist_tx = np.linalg.norm(diff_tx, axis=1)
        v_rad = (diff_tx / dist_tx[:, None]) @ (vel_plat[k] - v_f)
        t_shift = (-g["FC"] * (2 * v_rad / g["C"])) / g["K_RATE"]
        tau_a = 2 * dist_tx / g["C"]
        pos_rx = pos_plat[k][None, :] + vel_plat[k][None, :] * tau_a[:, None]
        g_rx = gk + v_f[None, :] * tau_a[:, None]
        tau = (dist_tx + np.linalg.norm(g_rx - pos_rx, axis=1)) / g["C"]
        idx_norm = (2 * (((tau - t_start + t_shift) * g["FS"]) / num_samples) - 1).astype(np.float32)
        samp = grid_sample_row_f32(rc_re[k], idx_norm).astype(np.float64) + 1j * grid_sample_row_f32(rc_im[k], idx_norm).astype(np.float64)
        img += samp * np.exp(1j * (2 * np.pi * g["FC"] * tau))
    return img.reshape(ny, nx)
