"""Point-scatterer vehicle models: drop-in for the reference's ``vehicle_targets`` module.

Same generator names, same ``(center_pos, name_prefix)`` arguments, same return type
-- a list of ``{'position': [x, y, z], 'rcs': float, 'name': str}`` dicts -- and the same
scatterer coordinates / RCS values, point for point (vehicle_targets.py:3-141; checked
against the reference module by tests/golden/vehicle_targets.npz).  The tables below are
laid out as data (offset triples per part) rather than the reference's inline loops.
"""
from __future__ import annotations

import numpy as np


def create_point_target(x, y, z, rcs, name=""):
    """vehicle_targets.py:3-4."""
    return {"position": [x, y, z], "rcs": rcs, "name": name}


def _emit(center_pos, offsets, rcs, prefix):
    cx, cy, cz = center_pos
    return [create_point_target(cx + ox, cy + oy, cz + oz, rcs, f"{prefix}_pt{i}")
            for i, (ox, oy, oz) in enumerate(offsets)]


def _rect(half_l, half_w, z):
    return [(half_l, half_w, z), (half_l, -half_w, z), (-half_l, half_w, z), (-half_l, -half_w, z)]


def generate_car(center_pos=(0, 0, 0), name_prefix="Car"):
    """12 unit-RCS points: chassis and roof rectangles, bumpers, door mid-points
    (vehicle_targets.py:6-41)."""
    length, width = 4.5, 1.8
    pts = (_rect(length / 2, width / 2, 0.5)
           + _rect(2.0 / 2, 1.4 / 2, 1.4)
           + [(length / 2, 0, 0.4), (-length / 2, 0, 0.4)]
           + [(0, width / 2, 0.9), (0, -width / 2, 0.9)])
    return _emit(center_pos, pts, 1.0, name_prefix)


def generate_tank(center_pos=(0, 0, 0), name_prefix="Tank"):
    """18 points of RCS 5: hull box, turret cross, gun barrel, hull mid-points
    (vehicle_targets.py:43-73)."""
    length, width, height = 8.0, 3.6, 1.5
    z_t, rad = 2.3, 1.5
    pts = (_rect(length / 2, width / 2, height) + _rect(length / 2, width / 2, 0.5)
           + [(0, 0, z_t), (rad, 0, z_t - 0.3), (-rad, 0, z_t - 0.3), (0, rad, z_t - 0.3), (0, -rad, z_t - 0.3)]
           + [(length / 2 + d, 0, z_t - 0.5) for d in (1.0, 3.0, 5.0)]
           + [(0, width / 2, 1.0), (0, -width / 2, 1.0)])
    return _emit(center_pos, pts, 5.0, name_prefix)


def generate_fighter_jet(center_pos=(0, 0, 0), name_prefix="Jet4Gen", rcs_scale=1.0):
    """13 points of RCS 10*scale: fuselage, wings, stabilisers (vehicle_targets.py:75-97)."""
    pts = ([(7.5, 0, 0), (5.0, 0, 1.0), (-6.0, 0, 1.0), (-7.0, 0, 0.5), (-6.0, 0, 2.5)]
           + [(0, 2.0, 0), (0, -2.0, 0), (-3.0, 5.0, 0), (-3.0, -5.0, 0), (-4.0, 2.5, 0), (-4.0, -2.5, 0)]
           + [(-6.5, 2.0, 0), (-6.5, -2.0, 0)])
    return _emit(center_pos, pts, 10.0 * rcs_scale, name_prefix)


def generate_f35(center_pos=(0, 0, 0), name_prefix="F35"):
    """Low-observable jet: the fighter at 1 % RCS (vehicle_targets.py:99-100)."""
    return generate_fighter_jet(center_pos, name_prefix, rcs_scale=0.01)


def generate_destroyer(center_pos=(0, 0, 0), name_prefix="Destroyer"):
    """154 m x 20 m hull: 5 x 3 grid at two heights (30 x 1000 m^2) plus bridge, mast,
    stack, bow and stern reflectors -- 35 points, 43 000 m^2 (vehicle_targets.py:102-141)."""
    cx, cy, cz = center_pos
    length, width = 154.0, 20.0
    out = []
    for x in np.linspace(-length / 2, length / 2, 5):
        for y in np.linspace(-width / 2, width / 2, 3):
            out.append(create_point_target(cx + x, cy + y, cz + 1, 1000.0, f"{name_prefix}_hull"))
            out.append(create_point_target(cx + x, cy + y, cz + 6, 1000.0, f"{name_prefix}_deck"))
    for dx, dz, rcs, tag in ((length * 0.2, 15, 5000.0, "bridge"), (length * 0.1, 25, 3000.0, "mast"),
                             (-length * 0.1, 12, 3000.0, "stack"), (length / 2.0 + 10.0, 6, 1000.0, "bow"),
                             (-length / 2.0 - 5.0, 6, 1000.0, "stern")):
        out.append(create_point_target(cx + dx, cy, cz + dz, rcs, f"{name_prefix}_{tag}"))
    return out


def targets_to_arrays(targets):
    """list-of-dict scatterers -> (pos[T,3] f64, rcs[T] f64), as every reference engine
    does on entry (sar_ati_dcpa_sim_csa.py:123-124)."""
    pos = np.array([t["position"] for t in targets], dtype=np.float64).reshape(-1, 3)
    rcs = np.array([t["rcs"] for t in targets], dtype=np.float64).reshape(-1)
    return pos, rcs


def arrays_to_targets(pos, rcs, name_prefix="pt"):
    return [{"position": np.asarray(p, dtype=np.float64), "rcs": float(r), "name": f"{name_prefix}{i}"}
            for i, (p, r) in enumerate(zip(pos, rcs))]
