"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, all work enqueued on
torch's current stream through the C ABI (include/nis_sar.h).  torch is the buffer carrier only."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import NisError


def _dev_index(device) -> int:
    d = torch.device(device)
    if d.type != "cuda":
        raise NisError(f"nis_sar runs on CUDA devices only (got {d}); there is no CPU fallback")
    return torch.cuda.current_device() if d.index is None else d.index


def _stream_ptr(dev_index: int) -> int:
    return torch.cuda.current_stream(dev_index).cuda_stream


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f64(x, device):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device, non_blocking=True)


# ------------------------------------------------------------------------------------- echo
MAX_PULSES_PER_LAUNCH = 65535


def fast_time_axis(t_start: float, n_samples: int, fs: float) -> np.ndarray:
    """The reference's receive-window grid: t_start + linspace(0, S/fs, S)
    (sar_ati_dcpa_sim_csa.py:113-114) -- S points including the end point."""
    return t_start + np.linspace(0.0, n_samples / fs, n_samples)


def chunk_hint(*_args, **_kw) -> int:
    """CTA chunk width of K1 (nis_echo_params.samples_per_thread): 0 lets the library choose.  Measured on
    B200 the narrow chunk (8) never wins -- the per-sample setup doubles -- so it stays an ABI knob only."""
    return 0


def echo_accumulate(pos0, vel, rcs, pos_tx, pos_rx, t_slow, *, c, fc, k_rate, t_p, t_start, fs, n_samples,
                    device="cuda", out=None, accumulate=False, pulse_range=None, spotlight=None):
    """K1.  Inputs are numpy / torch fp64 arrays (host or device); returns raw[P, S] complex64 on
    ``device``.  ``vel`` is one xyz triple or a [T,3] array; ``pos_rx`` None selects the monostatic
    delay 2|p - p_tx|/c.  ``pulse_range=(p0, p1)`` restricts the rows that are computed (the pulse
    block of one rank); rows outside are left untouched.  ``spotlight=(vel_sat[P,3], pi l_ant / lambda, t_fast[S])``
    ``accumulate``: False overwrite, True add, "atomic" add with device-scope-free reductions (``out`` may be another GPU's
    buffer mapped through ``nis_sar.dist.share_tensor``: the partial echoes of several ranks meet in one buffer).
    selects the run_physics_spotlight model (sar_batch_sim.py:83-169): start-stop corrected delay, sinc^2 pattern,
    amplitude rcs (not its square root), chirp centred on the delay, sample times t_start + n / fs."""
    di = _dev_index(device)
    dev = torch.device("cuda", di)
    lib = _lib.load()
    with torch.cuda.device(di):
        ctx = _lib.context(di)
        pos0_d = _f64(np.asarray(pos0).reshape(-1, 3) if not torch.is_tensor(pos0) else pos0.cpu().numpy().reshape(-1, 3), dev)
        T = pos0_d.shape[0]
        vel_np = np.asarray(vel.cpu().numpy() if torch.is_tensor(vel) else vel, dtype=np.float64)
        per_target = int(vel_np.size == 3 * T and vel_np.ndim == 2 and T > 1)
        vel_d = _f64(vel_np.reshape(-1), dev)
        rcs_np = np.asarray(rcs.cpu().numpy() if torch.is_tensor(rcs) else rcs, dtype=np.float64).reshape(-1)
        amp_d = _f64(rcs_np if spotlight is not None else np.sqrt(rcs_np), dev)
        if amp_d.shape[0] != T:
            raise NisError(f"echo_accumulate: {T} positions but {amp_d.shape[0]} rcs values")
        ptx_d = _f64(np.asarray(pos_tx).reshape(-1, 3), dev)
        P = ptx_d.shape[0]
        prx_d = None if pos_rx is None else _f64(np.asarray(pos_rx).reshape(-1, 3), dev)
        ts_d = _f64(np.asarray(t_slow).reshape(-1), dev)
        if ts_d.shape[0] != P or (prx_d is not None and prx_d.shape[0] != P):
            raise NisError("echo_accumulate: pos_tx / pos_rx / t_slow disagree on the number of pulses")
        S = int(n_samples)
        t_fast = fast_time_axis(t_start, S, fs) if spotlight is None else np.asarray(spotlight[2], dtype=np.float64)
        tf_d = _f64(t_fast, dev)
        if spotlight is not None:
            prx_d = _f64(np.asarray(spotlight[0], dtype=np.float64).reshape(-1, 3), dev)   # platform velocity per pulse
        if out is None:
            out = torch.zeros((P, S), dtype=torch.complex64, device=dev)
            accumulate = False if pulse_range is None else accumulate
        elif out.shape != (P, S) or out.dtype != torch.complex64 or not out.is_contiguous():
            raise NisError("echo_accumulate: out must be a contiguous complex64 [P, S] tensor")
        acc_mode = 2 if accumulate == "atomic" else (1 if accumulate else 0)
        p0, p1 = (0, P) if pulse_range is None else pulse_range
        prm = _lib.EchoParams(c=c, fc=fc, k_rate=k_rate, t_p=t_p, t_start=float(t_fast[0]),
                              dt_fast=((S / fs) / (S - 1) if S > 1 else 1.0 / fs) if spotlight is None else 1.0 / fs,
                              per_target_velocity=per_target,
                              samples_per_thread=chunk_hint(pos0_d, ptx_d, t_fast, t_p, c,
                                                            prx_d if spotlight is None else None))
        for q0 in range(p0, p1, MAX_PULSES_PER_LAUNCH):
            q1 = min(p1, q0 + MAX_PULSES_PER_LAUNCH)
            if spotlight is not None:
                rc = lib.nis_echo_spotlight(ctx, C.byref(prm), _ptr(pos0_d), _ptr(vel_d), _ptr(amp_d), _ptr(ptx_d),
                                            _ptr(prx_d), _ptr(ts_d), _ptr(tf_d), T, q0, q1, S, float(spotlight[1]),
                                            _ptr(out), acc_mode, C.c_void_p(_stream_ptr(di)))
                _lib.check(rc, "nis_echo_spotlight")
                continue
            rc = lib.nis_echo_accumulate(ctx, C.byref(prm), _ptr(pos0_d), _ptr(vel_d), _ptr(amp_d), _ptr(ptx_d),
                                         _ptr(prx_d), _ptr(ts_d), _ptr(tf_d), T, q0, q1, S, _ptr(out),
                                         acc_mode, C.c_void_p(_stream_ptr(di)))
            _lib.check(rc, "nis_echo_accumulate")
        # the fp64 input tensors may be freed right away: torch's allocator only reuses their memory
        # for later work on this same stream
    return out


# -------------------------------------------------------------------------------------- CSA
class CsaPlan:
    """K2 plan: twiddles, fp64-derived phase coefficients and workspace for one (n_az, n_rg) and one
    parameter set of ``sar_focus_csa`` (sar_ati_dcpa_sim_csa.py:202)."""

    def __init__(self, n_az, n_rg, *, lam, kr, fs, prf, vr, r_ref, t_start, c=299792458.0, device="cuda"):
        self.di = _dev_index(device)
        self.n_az, self.n_rg = int(n_az), int(n_rg)
        lib = _lib.load()
        prm = _lib.CsaParams(c=c, lambda_=lam, kr=kr, fs=fs, prf=prf, vr=vr, r_ref=r_ref, t_start=t_start)
        h = C.c_void_p()
        with torch.cuda.device(self.di):
            _lib.check(lib.nis_csa_plan_create(_lib.context(self.di), self.n_az, self.n_rg, C.byref(prm), C.byref(h)),
                       "nis_csa_plan_create")
        self._h = h
        self.key = (self.n_az, self.n_rg, lam, kr, fs, prf, vr, r_ref, t_start, c, self.di)

    @staticmethod
    def supported(n_az, n_rg) -> bool:
        return _lib.load().nis_csa_size_class(int(n_az), int(n_rg)) != 0

    def axes(self):
        ra = np.empty(self.n_rg, dtype=np.float64)
        ca = np.empty(self.n_az, dtype=np.float64)
        _lib.check(_lib.load().nis_csa_axes(self._h, ra.ctypes.data_as(C.c_void_p), ca.ctypes.data_as(C.c_void_p)),
                   "nis_csa_axes")
        return ra, ca

    def focus(self, phist, out=None, max_sq=None):
        """phist: complex64 CUDA tensor [n_az, n_rg] (rows may be strided: a ``raw[1:]`` view is fine).
        Returns slc [n_rg, n_az] complex64 -- the array the reference returns as ``img.T``.  ``max_sq``: optional
        1-element float64 CUDA tensor that receives max |slc|^2 (exact, for ``gmti_fused``; reset by the library)."""
        if phist.dtype != torch.complex64 or phist.dim() != 2 or phist.stride(1) != 1:
            raise NisError("CsaPlan.focus: phist must be a complex64 [n_az, n_rg] tensor with unit column stride")
        if tuple(phist.shape) != (self.n_az, self.n_rg):
            raise NisError(f"CsaPlan.focus: plan is {self.n_az}x{self.n_rg}, got {tuple(phist.shape)}")
        if out is None:
            out = torch.empty((self.n_rg, self.n_az), dtype=torch.complex64, device=phist.device)
        with torch.cuda.device(self.di):
            if max_sq is not None and (max_sq.dtype != torch.float64 or max_sq.numel() != 1):
                raise NisError("CsaPlan.focus: max_sq must be a 1-element float64 tensor")
            rc = _lib.load().nis_csa_focus(self._h, _ptr(phist), phist.stride(0), _ptr(out), _ptr(max_sq),
                                           C.c_void_p(_stream_ptr(self.di)))
        _lib.check(rc, "nis_csa_focus")
        return out

    STAGES = ("az_outer_fwd", "az_inner_fwd", "range", "az_inner_inv", "az_outer_inv")

    def set_profiling(self, enable=True):
        _lib.check(_lib.load().nis_csa_plan_set_profiling(self._h, 1 if enable else 0), "nis_csa_plan_set_profiling")

    def stage_times(self, calls_back=0):
        """Milliseconds spent in each of the five kernels of the focus call ``calls_back`` calls ago."""
        ms = (C.c_float * 5)()
        _lib.check(_lib.load().nis_csa_stage_times(self._h, int(calls_back), C.cast(ms, C.c_void_p)), "nis_csa_stage_times")
        return dict(zip(self.STAGES, (float(x) for x in ms)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().nis_csa_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plan_cache: dict = {}


def cached_plan(n_az, n_rg, **kw) -> CsaPlan:
    di = _dev_index(kw.get("device", "cuda"))
    key = (int(n_az), int(n_rg), kw["lam"], kw["kr"], kw["fs"], kw["prf"], kw["vr"], kw["r_ref"], kw["t_start"],
           kw.get("c", 299792458.0), di)
    pl = _plan_cache.get(key)
    if pl is None:
        if len(_plan_cache) >= 4:          # plans own an n_az*n_rg workspace: keep only a few
            _plan_cache.pop(next(iter(_plan_cache))).close()
        pl = CsaPlan(n_az, n_rg, **kw)
        _plan_cache[key] = pl
    return pl


# -------------------------------------------------------------------------------------- RDA
class RdaPlan:
    """Range-Doppler focusing plan for one (n_pulses, n_ranges) and one parameter set of ``sar_focus_rda``
    (sar_satellite_sim.py:356): matched-filter spectrum, azimuth weights, per-Doppler RCMC / azimuth-compression
    coefficients, azimuth engine and workspace.  Arrays are pulse-major [n_pulses, n_ranges]."""

    EXPORTS = ("phist_compressed", "range_doppler", "range_doppler_rcmc", "range_doppler_filtered")

    def __init__(self, n_pulses, n_ranges, *, lam, t_p, kr, fs, prf, vr, range_grp, c=299792458.0, device="cuda"):
        self.di = _dev_index(device)
        self.n_pulses, self.n_ranges = int(n_pulses), int(n_ranges)
        prm = _lib.RdaParams(c=c, lambda_=lam, t_p=t_p, kr=kr, fs=fs, prf=prf, vr=vr, range_grp=range_grp)
        h = C.c_void_p()
        with torch.cuda.device(self.di):
            _lib.check(_lib.load().nis_rda_plan_create(_lib.context(self.di), self.n_pulses, self.n_ranges, C.byref(prm),
                                                      C.byref(h)), "nis_rda_plan_create")
        self._h = h

    @staticmethod
    def supported(n_pulses, n_ranges, *, lam, t_p, kr, fs, prf, vr, range_grp, c=299792458.0) -> bool:
        prm = _lib.RdaParams(c=c, lambda_=lam, t_p=t_p, kr=kr, fs=fs, prf=prf, vr=vr, range_grp=range_grp)
        return _lib.load().nis_rda_supported(int(n_pulses), int(n_ranges), C.byref(prm)) != 0

    def axes(self):
        """(range_axis_centered[n_ranges], cross_range[n_pulses], doppler_freq[n_pulses])"""
        ra = np.empty(self.n_ranges, dtype=np.float64)
        ca = np.empty(self.n_pulses, dtype=np.float64)
        da = np.empty(self.n_pulses, dtype=np.float64)
        _lib.check(_lib.load().nis_rda_axes(self._h, ra.ctypes.data_as(C.c_void_p), ca.ctypes.data_as(C.c_void_p),
                                            da.ctypes.data_as(C.c_void_p)), "nis_rda_axes")
        return ra, ca, da

    def focus(self, phist, want=()):
        """phist: complex64 CUDA tensor [n_pulses, n_ranges] (unit column stride).  Returns a dict with ``image_mag``
        (float32 [n_pulses, n_ranges] = the reference's ``sar_image_mag.T``) and the requested exports (complex64,
        pulse- / Doppler-major, i.e. the transposes of the reference's [num_ranges, num_pulses] arrays)."""
        if phist.dtype != torch.complex64 or phist.dim() != 2 or phist.stride(1) != 1:
            raise NisError("RdaPlan.focus: phist must be a complex64 [n_pulses, n_ranges] tensor with unit column stride")
        if tuple(phist.shape) != (self.n_pulses, self.n_ranges):
            raise NisError(f"RdaPlan.focus: plan is {self.n_pulses}x{self.n_ranges}, got {tuple(phist.shape)}")
        for w in want:
            if w not in self.EXPORTS:
                raise NisError(f"RdaPlan.focus: unknown export {w!r}")
        shape = (self.n_pulses, self.n_ranges)
        out = {"image_mag": torch.empty(shape, dtype=torch.float32, device=phist.device)}
        for w in want:
            out[w] = torch.empty(shape, dtype=torch.complex64, device=phist.device)
        with torch.cuda.device(self.di):
            rc = _lib.load().nis_rda_focus(self._h, _ptr(phist), phist.stride(0), _ptr(out["image_mag"]),
                                           *[_ptr(out.get(w)) for w in self.EXPORTS], C.c_void_p(_stream_ptr(self.di)))
        _lib.check(rc, "nis_rda_focus")
        return out

    def close(self):
        if self._h is not None:
            _lib.load().nis_rda_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_rda_plan_cache: dict = {}


def cached_rda_plan(n_pulses, n_ranges, **kw) -> RdaPlan:
    di = _dev_index(kw.get("device", "cuda"))
    key = (int(n_pulses), int(n_ranges), kw["lam"], kw["t_p"], kw["kr"], kw["fs"], kw["prf"], kw["vr"], kw["range_grp"],
           kw.get("c", 299792458.0), di)
    pl = _rda_plan_cache.get(key)
    if pl is None:
        if len(_rda_plan_cache) >= 2:
            _rda_plan_cache.pop(next(iter(_rda_plan_cache))).close()
        pl = RdaPlan(n_pulses, n_ranges, **kw)
        _rda_plan_cache[key] = pl
    return pl


# ------------------------------------------------------------------------------------- TDBP
class TdbpPlan:
    """Time-domain backprojection plan (tdbp_gpu, sar_batch_sim.py:171-238): reference-chirp block spectrum, pixel axes."""

    def __init__(self, *, c, fc, k_rate, t_p, fs, t_start, n_samples, scene_size, nx=512, ny=512, device="cuda"):
        self.di = _dev_index(device)
        self.n_samples, self.nx, self.ny = int(n_samples), int(nx), int(ny)
        prm = _lib.TdbpParams(c=c, fc=fc, k_rate=k_rate, t_p=t_p, fs=fs, t_start=t_start, scene_size=scene_size,
                              n_samples=self.n_samples, nx=self.nx, ny=self.ny, reserved=0)
        h = C.c_void_p()
        with torch.cuda.device(self.di):
            _lib.check(_lib.load().nis_tdbp_plan_create(_lib.context(self.di), C.byref(prm), C.byref(h)), "nis_tdbp_plan_create")
        self._h = h

    def range_compress(self, raw):
        """raw: complex64 CUDA [P, n_samples] -> range-compressed pulses, same shape."""
        if raw.dtype != torch.complex64 or raw.dim() != 2 or raw.stride(1) != 1 or raw.shape[1] != self.n_samples:
            raise NisError("TdbpPlan.range_compress: raw must be complex64 [P, n_samples] with unit column stride")
        rc_t = torch.empty((raw.shape[0], self.n_samples), dtype=torch.complex64, device=raw.device)
        with torch.cuda.device(self.di):
            _lib.check(_lib.load().nis_tdbp_range_compress(self._h, _ptr(raw), raw.stride(0), raw.shape[0], _ptr(rc_t),
                                                          C.c_void_p(_stream_ptr(self.di))), "nis_tdbp_range_compress")
        return rc_t

    def backproject(self, rc_t, pos_plat, vel_plat, t_pulses, vel_focus, pulse_range=None, out=None, accumulate=False):
        """Sum pulses ``pulse_range`` (default: all) of the range-compressed CPI into the [ny, nx] complex128 image."""
        dev = rc_t.device
        P = rc_t.shape[0]
        pos_d, vel_d = _f64(np.asarray(pos_plat).reshape(-1, 3), dev), _f64(np.asarray(vel_plat).reshape(-1, 3), dev)
        t_np = np.asarray(t_pulses, dtype=np.float64).reshape(-1)
        tp_d = _f64(t_np, dev)
        if pos_d.shape[0] != P or vel_d.shape[0] != P or t_np.shape[0] != P:
            raise NisError("TdbpPlan.backproject: platform arrays disagree with the number of pulses")
        if out is None:
            out = torch.empty((self.ny, self.nx), dtype=torch.complex128, device=dev)
            accumulate = False
        p0, p1 = (0, P) if pulse_range is None else pulse_range
        vf = (C.c_double * 3)(*[float(v) for v in np.asarray(vel_focus, dtype=np.float64).reshape(3)])
        with torch.cuda.device(self.di):
            rc = _lib.load().nis_tdbp_backproject(self._h, _ptr(rc_t), _ptr(pos_d), _ptr(vel_d), _ptr(tp_d), P, int(p0), int(p1),
                                                  float(np.mean(t_np)), C.cast(vf, C.c_void_p), _ptr(out),
                                                  1 if accumulate else 0, C.c_void_p(_stream_ptr(self.di)))
        _lib.check(rc, "nis_tdbp_backproject")
        return out

    def close(self):
        if self._h is not None:
            _lib.load().nis_tdbp_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------ noise
def add_noise(x, snr_db, scr_db=10.0, k_nu=1.0, seed=0, ref_power=None, accumulate=True):
    """Thermal noise + K-distributed clutter into the complex64 CUDA tensor ``x`` in place (add_ocean_noise,
    sar_satellite_sim.py:331-344).  ``ref_power``: None = the mean power of ``x`` itself, "max" = its peak power
    (sar_batch_sim.py:317) -- both reduced on the device and consumed there, no host synchronisation --, or a number.
    ``accumulate`` False overwrites ``x`` with the noise alone (generate_noise_tensor, sar_batch_sim.py:65-81)."""
    if x.dtype != torch.complex64 or not x.is_contiguous():
        raise NisError("add_noise: x must be a contiguous complex64 CUDA tensor")
    di = x.device.index
    lib = _lib.load()
    n = x.numel()
    with torch.cuda.device(di):
        ctx = _lib.context(di)
        st = C.c_void_p(_stream_ptr(di))
        pdev, pval = None, 0.0
        if ref_power is None:
            pdev, pval = torch.empty(1, dtype=torch.float64, device=x.device), 1.0 / max(n, 1)
            _lib.check(lib.nis_power_sum(ctx, _ptr(x), n, _ptr(pdev), st), "nis_power_sum")
        elif isinstance(ref_power, str) and ref_power == "max":
            pdev, pval = torch.empty(1, dtype=torch.float64, device=x.device), 1.0
            _lib.check(lib.nis_power_max(ctx, _ptr(x), n, _ptr(pdev), st), "nis_power_max")
        elif torch.is_tensor(ref_power):
            pdev, pval = ref_power.to(device=x.device, dtype=torch.float64).reshape(1).contiguous(), 1.0
        else:
            pval = float(ref_power)
        _lib.check(lib.nis_noise_add(ctx, _ptr(x), n, _ptr(pdev), pval, float(snr_db), float(scr_db), float(k_nu),
                                     int(seed) & 0xFFFFFFFFFFFFFFFF, 1 if accumulate else 0, st), "nis_noise_add")
    return x


def peak_power(x):
    """max |x|^2 of a complex64 CUDA tensor as a 1-element float64 CUDA tensor (sar_batch_sim.py:317)."""
    di = x.device.index
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(di):
        _lib.check(_lib.load().nis_power_max(_lib.context(di), _ptr(x.contiguous()), x.numel(), _ptr(out),
                                             C.c_void_p(_stream_ptr(di))), "nis_power_max")
    return out


# ------------------------------------------------------------------------------------- GMTI
GMTI_PRODUCTS = ("ati_interf", "ati_phase", "dpca_diff", "dpca_mag", "slc1_mag", "mag_mask", "ati_phase_masked")


_GMTI_DTYPES = {"ati_interf": torch.complex64, "ati_phase": torch.float32, "dpca_diff": torch.complex64,
                "dpca_mag": torch.float32, "slc1_mag": torch.float32, "mag_mask": torch.uint8, "ati_phase_masked": torch.float32}


class GmtiBuffers:
    """Pre-allocated outputs of ``gmti_fused`` for one frame shape: the requested product maps, the detection list, the
    library's workspace and the 16-byte result record.  A VideoSAR loop that forms a pair per frame passes the same
    object every time (``gmti_fused(..., buffers=b, lazy=True)``): nothing is allocated per frame.  One object per
    stream -- two calls that may run concurrently must not share it."""

    def __init__(self, shape, want=GMTI_PRODUCTS, det_cap=None, device="cuda"):
        di = _dev_index(device)
        dev = torch.device("cuda", di)
        self.shape = tuple(int(v) for v in shape)
        n = int(np.prod(self.shape))
        self.want = tuple(want)
        self.products = {k: torch.empty(self.shape, dtype=_GMTI_DTYPES[k], device=dev) for k in self.want}
        self.cap = n if det_cap is None else int(det_cap)
        self.det = torch.empty((max(self.cap, 1),), dtype=torch.int32, device=dev)
        self.ws_bytes = int(_lib.load().nis_gmti_workspace_bytes(n))
        self.ws = torch.empty(((self.ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
        self.result = torch.empty((2,), dtype=torch.int64, device=dev).view(torch.uint8)


def gmti_fused(slc1, slc2, thresh_frac=0.05, cal_phase=0.0, want=GMTI_PRODUCTS, det_cap=None, max_sq=None,
               lazy=False, buffers: "GmtiBuffers | None" = None):
    """K3.  slc1/slc2: complex64 CUDA tensors of one shape.  Returns a dict with the requested
    product tensors (``max_sq``: optional 1-element float64 CUDA tensor filled by ``CsaPlan.focus(..., max_sq=)``
    for slc1 -- skips the max pass) plus ``det_idx`` (uint32 -> int64 tensor of flat indices, ascending),
    ``det_count``, ``peak_idx``, ``max_mag`` (python scalars; reading them synchronises).  ``lazy=True`` skips
    that read-back and returns the raw device buffers (``det_idx_raw``, ``result_dev``) instead.  ``buffers``: a
    ``GmtiBuffers`` of the same shape -- its tensors are written and returned, ``want`` / ``det_cap`` come from it."""
    if slc1.dtype != torch.complex64 or slc2.dtype != torch.complex64 or slc1.shape != slc2.shape:
        raise NisError("gmti_fused: slc1 and slc2 must be complex64 tensors of the same shape")
    if not (slc1.is_contiguous() and slc2.is_contiguous()):
        raise NisError("gmti_fused: SLCs must be contiguous")
    di = _dev_index(slc1.device)
    dev = slc1.device
    n = slc1.numel()
    shape = tuple(slc1.shape)
    outs = {}
    if buffers is not None:
        if buffers.shape != shape or buffers.det.device != dev:
            raise NisError(f"gmti_fused: buffers are for shape {buffers.shape} on {buffers.det.device}, got {shape} on {dev}")
        want, det_cap = buffers.want, buffers.cap
    with torch.cuda.device(di):
        def alloc(name, dtype):
            if name in want:
                outs[name] = buffers.products[name] if buffers is not None else torch.empty(shape, dtype=dtype, device=dev)
                return outs[name]
            return None
        interf = alloc("ati_interf", torch.complex64)
        phase = alloc("ati_phase", torch.float32)
        diff = alloc("dpca_diff", torch.complex64)
        dmag = alloc("dpca_mag", torch.float32)
        mag1 = alloc("slc1_mag", torch.float32)
        mask = alloc("mag_mask", torch.uint8)
        pmask = alloc("ati_phase_masked", torch.float32)
        cap = n if det_cap is None else int(det_cap)
        lib = _lib.load()
        if buffers is not None:
            det, ws_bytes, ws, res = buffers.det, buffers.ws_bytes, buffers.ws, buffers.result
        else:
            # per-call workspace and result record from torch's stream-ordered allocator (nothing is zero-filled here:
            # the library writes both); concurrent calls on other streams get their own
            det = torch.empty((max(cap, 1),), dtype=torch.int32, device=dev)
            ws_bytes = int(lib.nis_gmti_workspace_bytes(n))
            ws = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
            res = torch.empty((2,), dtype=torch.int64, device=dev).view(torch.uint8)
        rc = lib.nis_gmti_fused(_lib.context(di), _ptr(slc1), _ptr(slc2), n, float(thresh_frac),
                                float(cal_phase), _ptr(interf), _ptr(phase), _ptr(diff), _ptr(dmag),
                                _ptr(mag1), _ptr(mask), _ptr(pmask), _ptr(det), cap, _ptr(max_sq), _ptr(ws), ws_bytes,
                                _ptr(res), C.c_void_p(_stream_ptr(di)))
        _lib.check(rc, "nis_gmti_fused")
        if lazy:            # no host synchronisation: the caller reads the 16-byte record / index list later
            if "mag_mask" in outs:
                outs["mag_mask"] = outs["mag_mask"].view(torch.bool)
            outs["det_idx_raw"] = det
            outs["result_dev"] = res
            return outs
        raw = res.cpu().numpy().tobytes()          # synchronises: the 16-byte result record
    r = _lib.GmtiResult.from_buffer_copy(raw)
    k = min(int(r.det_count), cap)
    if "mag_mask" in outs:
        outs["mag_mask"] = outs["mag_mask"].view(torch.bool)
    outs["det_idx"] = det[:k].to(torch.int64) & 0xFFFFFFFF
    outs["det_count"] = int(r.det_count)
    outs["peak_idx"] = int(r.peak_idx)
    outs["max_mag"] = math.sqrt(r.max_mag_sq)
    return outs


def balance_phase(slc1, slc2) -> float:
    """Viewer auto-balance angle(mean(slc1 conj(slc2))) (sar_ati_dcpa_viewer_csa.py:249-250)."""
    di = _dev_index(slc1.device)
    with torch.cuda.device(di):
        acc = torch.zeros(2, dtype=torch.float64, device=slc1.device)
        rc = _lib.load().nis_gmti_balance_sum(_lib.context(di), _ptr(slc1.contiguous()), _ptr(slc2.contiguous()),
                                              slc1.numel(), _ptr(acc), C.c_void_p(_stream_ptr(di)))
        _lib.check(rc, "nis_gmti_balance_sum")
        re, im = acc.cpu().tolist()
    return math.atan2(im, re)


# ---------------------------------------------------------------------------------- formats
def narrow_c128(x128, out=None):
    di = _dev_index(x128.device)
    if out is None:
        out = torch.empty(x128.shape, dtype=torch.complex64, device=x128.device)
    with torch.cuda.device(di):
        _lib.check(_lib.load().nis_narrow_c128_to_c32(_lib.context(di), _ptr(x128), _ptr(out), x128.numel(),
                                                      C.c_void_p(_stream_ptr(di))), "nis_narrow_c128_to_c32")
    return out


def widen_c32(x64, out=None):
    di = _dev_index(x64.device)
    if out is None:
        out = torch.empty(x64.shape, dtype=torch.complex128, device=x64.device)
    with torch.cuda.device(di):
        _lib.check(_lib.load().nis_widen_c32_to_c128(_lib.context(di), _ptr(x64), _ptr(out), x64.numel(),
                                                     C.c_void_p(_stream_ptr(di))), "nis_widen_c32_to_c128")
    return out


def transpose_c32(x, out=None):
    di = _dev_index(x.device)
    rows, cols = x.shape
    if out is None:
        out = torch.empty((cols, rows), dtype=torch.complex64, device=x.device)
    with torch.cuda.device(di):
        _lib.check(_lib.load().nis_transpose_c32(_lib.context(di), _ptr(x), _ptr(out), rows, cols,
                                                 C.c_void_p(_stream_ptr(di))), "nis_transpose_c32")
    return out
