"""Radar / geometry constants the reference keeps as module globals, made explicit.

The reference functions read ``C, R0, FC, BW, T_p, FS`` from their module's global
scope (sar_ati_dcpa_sim_csa.py:111-115, :159-168); here they travel as a
``RadarParams`` value.  Presets reproduce the three parameter sets on the hot path.
"""
from __future__ import annotations

import dataclasses
import math

C_LIGHT = 299792458.0


@dataclasses.dataclass(frozen=True)
class RadarParams:
    C: float = C_LIGHT
    FC: float = 9.65e9          # carrier (Hz)
    BW: float = 500e6           # chirp bandwidth (Hz)
    T_p: float = 20e-6          # pulse width (s)
    FS: float = 600e6           # fast-time sample rate (Hz)
    PRF: float = 6000.0
    R0: float = 0.0             # scene-centre slant range (m)
    V_sat: float = 0.0          # platform speed (m/s)
    V_eff: float = 0.0          # effective focusing speed (m/s)
    R_sat: float = 0.0          # orbit radius (m); 0 for an airborne platform
    Re: float = 6371000.0
    gamma_rad: float = 0.0      # Earth-centre angle between target and platform at broadside
    window_s: float = 22e-6     # receive window length (s): S = int(window_s * FS)
    n_samples: int = 0          # explicit S (0 = derive from window_s as the reference does)

    @property
    def Lambda(self) -> float:
        return self.C / self.FC

    @property
    def k_rate(self) -> float:
        return self.BW / self.T_p

    @property
    def num_samples(self) -> int:
        return self.n_samples if self.n_samples > 0 else int(self.window_s * self.FS)

    @property
    def d_rx(self) -> float:
        """DPCA phase-centre separation: one PRI of platform motion x 2
        (sar_ati_dcpa_sim_csa.py:42)."""
        return 2 * self.V_sat / self.PRF

    @property
    def t_start_fast(self) -> float:
        """Opening of the receive window (sar_ati_dcpa_sim_csa.py:112)."""
        return (2 * self.R0 / self.C) - (self.T_p / 2) - 1e-6

    def as_globals(self) -> dict:
        """The dict the oracle / extracted reference functions take."""
        return {"C": self.C, "R0": self.R0, "FC": self.FC, "BW": self.BW, "T_p": self.T_p, "FS": self.FS}

    def replace(self, **kw) -> "RadarParams":
        return dataclasses.replace(self, **kw)


def spaceborne_preset(fs: float = 600e6, bw: float = 500e6, t_p: float = 20e-6,
                      window_s: float = 22e-6, prf: float = 6000.0) -> RadarParams:
    """350 km circular orbit, 45 deg look angle, X band -- the constants block shared by
    sar_ati_dcpa_sim_csa.py:18-38,:68 and sar_satellite_sim.py:23-60,:184."""
    Re = 6371000.0
    h = 350000.0
    R_sat = Re + h
    GM = 3.986004418e14
    V_sat = math.sqrt(GM / R_sat)
    look = math.radians(45.0)
    inc = math.asin((R_sat / Re) * math.sin(look))
    gamma = inc - look
    R0 = math.sqrt(Re ** 2 + R_sat ** 2 - 2 * Re * R_sat * math.cos(gamma))
    V_eff = V_sat * math.sqrt(Re / R_sat)
    return RadarParams(FC=9.65e9, BW=bw, T_p=t_p, FS=fs, PRF=prf, R0=R0, V_sat=V_sat, V_eff=V_eff,
                       R_sat=R_sat, Re=Re, gamma_rad=gamma, window_s=window_s)


def airborne_vehicle_preset() -> RadarParams:
    """20 km altitude, 150 m/s, 10 GHz, 300 MHz / 1 us chirp, 360 MHz sampling, 2048 samples
    (sar_vehicle_sim.py:22-38, :85-86, :168-170)."""
    Re = 6378137.0
    h = 20000.0
    look = math.radians(45.0)
    R0 = h / math.cos(look)
    return RadarParams(FC=10e9, BW=300e6, T_p=1.0e-6, FS=360e6, PRF=1.0 / 500e-6, R0=R0,
                       V_sat=150.0, V_eff=150.0, R_sat=0.0, Re=Re, gamma_rad=0.0,
                       window_s=2048 / 360e6, n_samples=2048)


def batch_spotlight_preset(fs: float = 600e6, bw: float = 500e6, t_p: float = 20e-6, prf: float = 5000.0) -> RadarParams:
    """The constants block of sar_batch_sim.py:12-38 (same orbit as the spaceborne preset, PRF 5 kHz)."""
    return spaceborne_preset(fs=fs, bw=bw, t_p=t_p, prf=prf)
