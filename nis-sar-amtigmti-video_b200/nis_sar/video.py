"""VideoSAR frame loop of sar_batch_sim.py (:289-326) on the device: a sliding coherent processing interval (CPI) over a
long pulse train; per frame spotlight echo synthesis -> noise at the peak echo power -> time-domain backprojection.
Frames are independent: with ``torch.distributed`` initialised every rank takes ``nis_sar.dist.frame_indices`` (round
robin) and no collective is needed -- the partition SURVEY.md section 8e lists for VideoSAR frames."""
from __future__ import annotations

import numpy as np

from . import api, dist as nd
from . import device as dev
from .params import RadarParams


def cpi_windows(total_pulses: int, step_pulses: int, cpi_pulses: int, num_frames: int):
    """(i0, i1) of every frame whose CPI fits the pulse train (sar_batch_sim.py:305-308)."""
    out = []
    for f in range(num_frames):
        i0 = f * step_pulses
        i1 = i0 + cpi_pulses
        if i1 > total_pulses:
            break
        out.append((i0, i1))
    return out


def batch_timeline(prm: RadarParams, duration=5.0, fps=10, cpi_s=0.5):
    """Pulse train of sar_batch_sim.py:243-264: (t_vec_all, step_pulses, cpi_pulses, num_frames)."""
    total = int(np.ceil(duration * prm.PRF))
    return (np.linspace(-duration / 2, duration / 2, total), int(prm.PRF / fps), int(np.ceil(cpi_s * prm.PRF)),
            int(duration * fps))


def render_frames(base_targets, t_vec_all, pos_sat_all, vel_sat_all, *, heading_deg, speed, l_ant, scene_size,
                  step_pulses, cpi_pulses, num_frames, focus_tgt=True, snr_db=None, nx=512, ny=512, seed=0,
                  params: RadarParams | None = None, device=None, frames=None, return_device=False):
    """Frames of one (vehicle, heading, algorithm) run (sar_batch_sim.py:300-326): dict frame index -> complex128 image
    [ny, nx].  ``focus_tgt`` True = mBP (pixels move with the target), False = StdBP.  ``snr_db`` None skips the noise
    (:317-318 otherwise: noise + clutter at the peak power of the CPI's echo).  ``frames``: the indices this caller
    renders (default: this rank's round-robin share when torch.distributed is initialised, else all)."""
    wins = cpi_windows(len(t_vec_all), step_pulses, cpi_pulses, num_frames)
    mine = list(frames) if frames is not None else list(nd.frame_indices(len(wins)))
    out = {}
    for f in mine:
        i0, i1 = wins[f]
        t_cpi, p_cpi, v_cpi = t_vec_all[i0:i1], pos_sat_all[i0:i1], vel_sat_all[i0:i1]
        raw, t_st, n_sp, v_tgt = api.run_physics_spotlight(base_targets, t_cpi, p_cpi, v_cpi, heading_deg, speed, l_ant,
                                                           params=params, device=device)
        if snr_db is not None:
            dev.add_noise(raw, snr_db, seed=seed + f, ref_power="max")
        vf = v_tgt if focus_tgt else np.zeros(3)
        out[f] = api.tdbp_gpu(raw, p_cpi, v_cpi, t_st, n_sp, vf, t_cpi, scene_size, nx=nx, ny=ny, params=params,
                              device=device, return_device=return_device)
    return out
