"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's own torch formulation of the bistatic echo engine.

``run_bistatic_physics_gpu`` (sar_ati_dcpa_sim_csa.py:106-181) is already a GPU program: a Python loop over pulses, each
iteration ~15 eager torch kernels over [T scatterers, S samples] complex128 temporaries.  This restatement keeps that
structure (same tensor shapes, dtypes and operation order) so that ``bench.py`` can time "the reference's GPU path" on the
B200 next to the hand-written kernel -- on a bounded pulse subset, as a reported baseline.  Nothing under
``nis-sar-amtigmti-video_b200/`` imports it.  Pinned against the numpy oracle (itself pinned to the reference's outputs) in
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np
import torch


@torch.no_grad()
def echo_bistatic_torch(pos0, rcs, t_vec, pos_tx, vel_tx, rx_offset, vel_target, g, n_samples=None, device="cpu"):
    """g: dict with C, R0, FC, BW, T_p, FS.  Returns (raw[P, S] complex128 torch tensor on ``device``, t_start_fast)."""
    device = torch.device(device)
    C, R0, FC, BW, T_p, FS = (g[k] for k in ("C", "R0", "FC", "BW", "T_p", "FS"))
    num_samples = int(22e-6 * FS) if n_samples is None else int(n_samples)           # :111
    t_start_fast = (2 * R0 / C) - (T_p / 2) - 1e-6                                     # :112
    t_fast_abs = t_start_fast + np.linspace(0, num_samples / FS, num_samples)          # :113-114
    k_rate = BW / T_p
    f64 = dict(device=device, dtype=torch.float64)
    pos_tx_t, vel_tx_t = torch.tensor(np.asarray(pos_tx), **f64), torch.tensor(np.asarray(vel_tx), **f64)
    t_pos_0_t = torch.tensor(np.asarray(pos0), **f64).view(-1, 3)
    amp = torch.sqrt(torch.tensor(np.asarray(rcs), **f64).view(-1, 1))                 # :169
    vel_target_t = torch.tensor(np.asarray(vel_target, dtype=float), **f64).view(1, 3)
    t_fast_t = torch.tensor(t_fast_abs, **f64).view(1, -1)
    raw_sig = torch.zeros((len(t_vec), num_samples), device=device, dtype=torch.complex128)
    for i in range(len(t_vec)):                                                         # :137
        p_tx = pos_tx_t[i].view(1, 3)
        v_tx = vel_tx_t[i].view(1, 3)
        p_rx = p_tx + (v_tx / torch.norm(v_tx)) * rx_offset                            # :145-148
        t_pos_curr = t_pos_0_t + vel_target_t * float(t_vec[i])                        # :151
        tau = (torch.norm(t_pos_curr - p_tx, dim=1) + torch.norm(t_pos_curr - p_rx, dim=1)) / C   # :154-159
        phase_base = -2.0 * np.pi * FC * tau
        t_local = t_fast_t - tau.view(-1, 1)
        mask = torch.abs(t_local - T_p / 2) <= (T_p / 2)                               # :166
        chirp = np.pi * k_rate * ((t_local - T_p / 2) ** 2)
        raw_sig[i] = torch.sum(amp * torch.exp(1j * (phase_base.view(-1, 1) + chirp)) * mask, dim=0)   # :171-176
    return raw_sig, t_start_fast
