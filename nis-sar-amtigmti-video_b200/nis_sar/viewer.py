"""Data layer of the interactive ATI / DPCA viewer (sar_ati_dcpa_viewer_csa.py) on the device.

The reference viewer loads ``slc1, slc2`` from the simulator's ``.npz`` (transposing them back to
[N_cross, N_range], :24-27), keeps seven derived product maps in a ``SARData`` object (:35-55) and, on every mode
change, zoom or pan, recomputes numpy statistics of the visible rectangle: mean / median / std / min / max (:117-139),
the DPCA cancellation ratio (:141-145) and the 99.9th-percentile colour limit (:147-151, :176-186); "Auto-Balance"
re-derives everything with ``cal_phase = angle(mean(slc1 conj(slc2)))`` (:245-253).  Here the maps live in HBM
(one fused pass), the statistics are reductions / radix selects on the visible rectangle, and only scalars -- or a map the
GUI actually draws -- cross PCIe.  Same names, same orientation: arrays are [N_cross, N_range] as in the viewer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import device as dev
from . import hostio

MODES = ("Ch1 Magnitude", "Ch1 Phase", "Ch2 Magnitude", "Ch2 Phase", "DPCA Magnitude", "DPCA Phase", "ATI Phase")


def _lerp(a, b, t):
    """numpy's percentile interpolation (lib/_function_base_impl.py:_lerp)."""
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


class SARData:
    """``SARData(s1, s2)`` of sar_ati_dcpa_viewer_csa.py:35-55.  ``s1``/``s2``: complex arrays in the viewer's
    orientation [N_cross, N_range] -- numpy (the ``data['slc1'].T`` views are consumed without a copy) or complex64 CUDA
    tensors.  ``get(mode)`` returns the float64 numpy map the reference's ``get`` returns; ``get_device(mode)`` the
    float32 CUDA tensor (storage orientation [N_range, N_cross])."""

    def __init__(self, s1, s2, device=None):
        self.device = torch.device(device or "cuda")
        self._s1 = self._upload(s1)
        self._s2 = self._upload(s2)
        if self._s1.shape != self._s2.shape:
            raise dev.NisError("SARData: the two channels differ in shape")
        self.n_range, self.n_cross = self._s1.shape
        self.cal_phase = 0.0
        self._maps = None
        self.compute_all()

    def _upload(self, a):
        if torch.is_tensor(a):
            t = a.transpose(0, 1)
            t = t if t.dtype == torch.complex64 else dev.narrow_c128(t.to(torch.complex128).contiguous())
            return t.contiguous().to(self.device)
        h = np.asarray(a).T                                     # storage orientation [N_range, N_cross]
        return hostio.to_device_c64(h, self.device)

    def compute_all(self):
        """All seven maps in one pass over the two images (:42-52)."""
        di = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if self._maps is None:
            self._maps = torch.empty((len(MODES), self.n_range, self.n_cross), dtype=torch.float32, device=self.device)
        with torch.cuda.device(di):
            rc = _lib.load().nis_viewer_products(_lib.context(di), dev._ptr(self._s1), dev._ptr(self._s2),
                                                 self._s1.numel(), float(self.cal_phase),
                                                 *[dev._ptr(self._maps[i]) for i in range(len(MODES))],
                                                 C.c_void_p(dev._stream_ptr(di)))
        _lib.check(rc, "nis_viewer_products")
        self._di = di

    def get_device(self, mode):
        return self._maps[MODES.index(mode)] if mode in MODES else None

    def get(self, mode):
        m = self.get_device(mode)
        return None if m is None else m.to(torch.float64).cpu().numpy().T

    def balance(self):
        """Auto-Balance (:245-253): cal_phase = angle(mean(slc1 conj(slc2))), then recompute."""
        self.cal_phase = dev.balance_phase(self._s1, self._s2)
        self.compute_all()
        return self.cal_phase

    # ------------------------------------------------------------------ statistics of the visible rectangle
    def _region(self, mode, c_indices, r_indices):
        m = self.get_device(mode)
        c0, c1 = (0, self.n_cross) if c_indices is None else (int(c_indices[0]), int(c_indices[-1]) + 1)
        r0, r1 = (0, self.n_range) if r_indices is None else (int(r_indices[0]), int(r_indices[-1]) + 1)
        if not (0 <= c0 < c1 <= self.n_cross and 0 <= r0 < r1 <= self.n_range):
            raise dev.NisError("SARData: empty or out-of-range region")
        return m[r0:r1, c0:c1], r1 - r0, c1 - c0

    def _select(self, view, rows, cols, ranks):
        out = torch.empty(len(ranks), dtype=torch.float32, device=self.device)
        arr = (C.c_uint64 * len(ranks))(*[int(r) for r in ranks])
        with torch.cuda.device(self._di):
            rc = _lib.load().nis_region_select(_lib.context(self._di), dev._ptr(view), view.stride(0), rows, cols,
                                               C.cast(arr, C.c_void_p), len(ranks), dev._ptr(out),
                                               C.c_void_p(dev._stream_ptr(self._di)))
        _lib.check(rc, "nis_region_select")
        return out.cpu().numpy().astype(np.float64)

    @staticmethod
    def _display(v, db):
        return 20 * np.log10(v + 1e-12) if db else v

    def percentile(self, mode, q, scale="Linear", c_indices=None, r_indices=None):
        """np.percentile(display_data, q) of the region (linear interpolation between order statistics)."""
        view, rows, cols = self._region(mode, c_indices, r_indices)
        n = rows * cols
        pos = (n - 1) * (q / 100.0)
        lo = int(np.floor(pos))
        hi = min(lo + 1, n - 1)
        v = self._display(self._select(view, rows, cols, [lo, hi]), scale == "dB" and "Phase" not in mode)
        return float(_lerp(v[0], v[1], pos - lo))

    def clim(self, mode, scale="dB", c_indices=None, r_indices=None):
        """Colour limits of update_plot / print_visible_stats (:147-151, :176-186)."""
        if "Phase" in mode:
            return -np.pi, np.pi
        vmax = self.percentile(mode, 99.9, scale, c_indices, r_indices)
        return (vmax - 60 if scale == "dB" else 0), vmax

    def visible_stats(self, mode, scale="dB", c_indices=None, r_indices=None):
        """What print_visible_stats prints (:117-145): mean, median, std, min, max of the displayed values in the
        rectangle (phase modes: radians; magnitude modes: dB or linear) and, for DPCA modes, the local cancellation ratio
        mean(|ch1|) / (mean(visible) + 1e-9) on the linear values."""
        view, rows, cols = self._region(mode, c_indices, r_indices)
        n = rows * cols
        db = scale == "dB" and "Phase" not in mode
        acc = torch.empty(4, dtype=torch.float64, device=self.device)
        lib, ctx, st = _lib.load(), _lib.context(self._di), C.c_void_p(dev._stream_ptr(self._di))
        with torch.cuda.device(self._di):
            _lib.check(lib.nis_region_stats(ctx, dev._ptr(view), view.stride(0), rows, cols, 1 if db else 0, dev._ptr(acc), st),
                       "nis_region_stats")
        med = self._display(self._select(view, rows, cols, [(n - 1) // 2, n // 2]), db)
        s, mn, mx, ssd = acc.cpu().numpy()
        out = {"mean": s / n, "median": 0.5 * (med[0] + med[1]), "std": float(np.sqrt(ssd / n)), "min": mn, "max": mx,
               "n": n}
        if "DPCA" in mode:
            ref, _, _ = self._region("Ch1 Magnitude", c_indices, r_indices)
            acc2 = torch.empty(4, dtype=torch.float64, device=self.device)
            lin = acc if not db else torch.empty(4, dtype=torch.float64, device=self.device)
            with torch.cuda.device(self._di):
                _lib.check(lib.nis_region_stats(ctx, dev._ptr(ref), ref.stride(0), rows, cols, 0, dev._ptr(acc2), st),
                           "nis_region_stats")
                if db:
                    _lib.check(lib.nis_region_stats(ctx, dev._ptr(view), view.stride(0), rows, cols, 0, dev._ptr(lin), st),
                               "nis_region_stats")
            out["cancellation_ratio"] = float((acc2[0].item() / n) / (lin[0].item() / n + 1e-9))
        return out


def load_npz(fname, device=None):
    """The viewer's loading step (:24-31): returns (SARData, range_axis, cross_range, extent)."""
    data = np.load(fname)
    rax, cax = data["range_axis"], data["cross_range"]
    return SARData(data["slc1"].T, data["slc2"].T, device=device), rax, cax, [rax[0], rax[-1], cax[0], cax[-1]]
