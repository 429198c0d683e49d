"""Seeded scene / trajectory builders for the hot-path workloads.

The reference builds these inline at module top level with unseeded ``np.random`` calls
(sar_ati_dcpa_sim_csa.py:47-66, :78-100), so its default scene is not reproducible; here
every random draw goes through ``np.random.default_rng(seed)``.
"""
from __future__ import annotations

import numpy as np

from .params import RadarParams, spaceborne_preset, airborne_vehicle_preset
from . import targets as _tg


def slow_time(prm: RadarParams, num_pulses: int, t_int: float | None = None) -> np.ndarray:
    """``t_vec = linspace(-T/2, T/2, P)`` (sar_ati_dcpa_sim_csa.py:46-48).  With ``t_int`` None
    the aperture is P/PRF so that the pulse spacing is the PRI."""
    t_int = num_pulses / prm.PRF if t_int is None else t_int
    return np.linspace(-t_int / 2, t_int / 2, num_pulses)


def orbit_trajectory(prm: RadarParams, t_vec: np.ndarray, along: str = "y"):
    """Circular-orbit platform state in scene coordinates (target at the origin, Earth centre at
    (0, 0, -Re)): P(t) = S0 cos(wt) + R_sat v sin(wt) - (0,0,Re), V(t) = V_sat v cos(wt) - S0 w sin(wt).
    ``along='y'``: S0 = (-R sin g, 0, R cos g), v = +y (sar_ati_dcpa_sim_csa.py:50-66);
    ``along='x'``: S0 = (0, -R sin g, R cos g), v = +x (sar_satellite_sim.py:130-172)."""
    w = prm.V_sat / prm.R_sat
    sg, cg = np.sin(prm.gamma_rad), np.cos(prm.gamma_rad)
    if along == "y":
        s0 = np.array([-prm.R_sat * sg, 0.0, prm.R_sat * cg])
        v_unit = np.array([0.0, 1.0, 0.0])
    else:
        s0 = np.array([0.0, -prm.R_sat * sg, prm.R_sat * cg])
        v_unit = np.array([1.0, 0.0, 0.0])
    wt = (w * np.asarray(t_vec, dtype=np.float64))[:, None]
    pos = s0[None, :] * np.cos(wt) + (prm.R_sat * v_unit)[None, :] * np.sin(wt) + np.array([0.0, 0.0, -prm.Re])
    vel = (prm.V_sat * v_unit)[None, :] * np.cos(wt) - (s0 * w)[None, :] * np.sin(wt)
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def straight_trajectory(prm: RadarParams, t_vec: np.ndarray):
    """Airborne straight-and-level pass: x = -R0 sin(45deg), y = V t, z = R0 cos(45deg)
    (sar_vehicle_sim.py:60-71)."""
    look = np.radians(45.0)
    pos = np.zeros((len(t_vec), 3))
    pos[:, 0] = -prm.R0 * np.sin(look)
    pos[:, 1] = prm.V_sat * np.asarray(t_vec)
    pos[:, 2] = prm.R0 * np.cos(look)
    vel = np.zeros_like(pos)
    vel[:, 1] = prm.V_sat
    return pos, vel


def ocean_clutter(seed: int, num_clutter: int = 5000, half_width: float = 3000.0, sigma0_db: float = 5.0):
    """Point-scatterer sea clutter: uniform over +-half_width, RCS ~ Exp(mean = area * sigma0 / n)
    (sar_ati_dcpa_sim_csa.py:78-100).  Returns (pos[T,3], rcs[T])."""
    rng = np.random.default_rng(seed)
    mean_rcs = ((2 * half_width) ** 2) * (10 ** (sigma0_db / 10.0)) / num_clutter
    pos = np.zeros((num_clutter, 3))
    pos[:, 0] = rng.uniform(-half_width, half_width, num_clutter)
    pos[:, 1] = rng.uniform(-half_width, half_width, num_clutter)
    rcs = rng.exponential(mean_rcs, num_clutter)
    return pos, rcs


def point_grid(n_side: int = 9, half_extent: float = 1000.0, rcs: float = 1.0):
    """Regular n x n grid of equal-RCS points on the ground plane (stripmap test scene)."""
    ax = np.linspace(-half_extent, half_extent, n_side)
    xx, yy = np.meshgrid(ax, ax, indexing="ij")
    pos = np.stack([xx.ravel(), yy.ravel(), np.zeros(n_side * n_side)], axis=1)
    return pos, np.full(n_side * n_side, rcs)


def dense_vehicle_scene(seed: int, num_scatterers: int, half_extent: float = 150.0):
    """Config 3: cars / tanks / destroyers tiled at seeded random centres inside the +-150 m
    scene of sar_vehicle_sim.py:48 until ``num_scatterers`` points exist."""
    rng = np.random.default_rng(seed)
    gens = (_tg.generate_car, _tg.generate_tank, _tg.generate_destroyer)
    pos_l, rcs_l, total = [], [], 0
    while total < num_scatterers:
        g = gens[int(rng.integers(0, len(gens)))]
        c = (float(rng.uniform(-half_extent, half_extent)), float(rng.uniform(-half_extent, half_extent)), 0.0)
        p, r = _tg.targets_to_arrays(g(center_pos=c))
        pos_l.append(p)
        rcs_l.append(r)
        total += len(r)
    return np.concatenate(pos_l)[:num_scatterers], np.concatenate(rcs_l)[:num_scatterers]


def ati_scene(seed: int = 0, num_pulses: int = 7200, num_clutter: int = 5000,
              prm: RadarParams | None = None, t_int: float | None = 1.2):
    """The default two-channel scene of sar_ati_dcpa_sim_csa.py: destroyer moving at
    (15, 0, 0) m/s + stationary clutter, receivers at -+d_rx/2 (:184-196).  ``num_pulses`` /
    ``num_clutter`` shrink it for tests; ``t_int=None`` keeps the pulse spacing at the PRI."""
    prm = spaceborne_preset() if prm is None else prm
    if t_int is not None and num_pulses != int(np.ceil(t_int * prm.PRF)):
        t_int = None
    t_vec = slow_time(prm, num_pulses, t_int)
    pos_tx, vel_tx = orbit_trajectory(prm, t_vec, along="y")
    ship_pos, ship_rcs = _tg.targets_to_arrays(_tg.generate_destroyer(center_pos=(0, 0, 0)))
    clut_pos, clut_rcs = ocean_clutter(seed, num_clutter) if num_clutter > 0 else (np.zeros((0, 3)), np.zeros(0))
    return {
        "prm": prm, "t_vec": t_vec, "pos_tx": pos_tx, "vel_tx": vel_tx,
        "ship_pos": ship_pos, "ship_rcs": ship_rcs, "ship_vel": np.array([15.0, 0.0, 0.0]),
        "clutter_pos": clut_pos, "clutter_rcs": clut_rcs, "clutter_vel": np.zeros(3),
        "rx_offsets": (-prm.d_rx / 2, prm.d_rx / 2),
    }


def hrws_scene(n_channels: int = 8, **kw):
    """Config 5 (HRWS-N fast-mover scene): the two-channel scene above with ``n_channels`` receive phase centres at
    (k - (N-1)/2) d_rx along the velocity vector, so that every adjacent pair (k, k+1) has the DPCA geometry of
    sar_ati_dcpa_sim_csa.py:42, :184-196 and is co-registered by the same one-pulse shift (:402-403)."""
    sc = ati_scene(**kw)
    d = sc["prm"].d_rx
    sc["rx_offsets"] = tuple((k - (n_channels - 1) / 2) * d for k in range(n_channels))
    return sc


def stripmap_scene(num_pulses: int, num_samples: int, n_side: int = 9, half_extent: float = 1000.0):
    """Config 2: sar_satellite_sim.py geometry (orbit along +x) with a point-target grid, the
    receive window sized so that S = ``num_samples`` at fs = 600 MHz."""
    prm = spaceborne_preset().replace(n_samples=num_samples, window_s=num_samples / 600e6)
    t_vec = slow_time(prm, num_pulses, None)
    pos_sat, vel_sat = orbit_trajectory(prm, t_vec, along="x")
    pos, rcs = point_grid(n_side, half_extent)
    return {"prm": prm, "t_vec": t_vec, "pos_sat": pos_sat, "vel_sat": vel_sat, "pos": pos, "rcs": rcs}


def vehicle_scene(seed: int, num_pulses: int = 32768, num_scatterers: int = 100000):
    """Config 3 geometry (sar_vehicle_sim.py:22-71)."""
    prm = airborne_vehicle_preset()
    t_int = num_pulses * 500e-6
    t_vec = np.linspace(-t_int / 2, t_int / 2, num_pulses)
    pos_plat, vel_plat = straight_trajectory(prm, t_vec)
    pos, rcs = dense_vehicle_scene(seed, num_scatterers)
    return {"prm": prm, "t_vec": t_vec, "pos_sat": pos_plat, "vel_sat": vel_plat, "pos": pos, "rcs": rcs}
