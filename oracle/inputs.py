"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by the fixture generator and the tests.

The large CSA fixtures (tests/golden/csa_large.npz) store only digests of the reference's output; the
input is regenerated from its seed wherever the test runs (numpy's PCG64 stream is platform independent).
"""
from __future__ import annotations

import numpy as np


def seeded_c64(shape, seed: int) -> np.ndarray:
    """complex64 [shape] with unit-variance Gaussian real / imaginary parts drawn as float32."""
    rng = np.random.default_rng(seed)
    re = rng.standard_normal(shape, dtype=np.float32)
    im = rng.standard_normal(shape, dtype=np.float32)
    out = np.empty(shape, dtype=np.complex64)
    out.real = re
    out.imag = im
    return out


def add_point_echoes(x: np.ndarray, seed: int, n_points: int = 6, gain: float = 40.0) -> np.ndarray:
    """A few strong single-sample impulses on top of the noise: after focusing they spread over the whole
    frame (the CSA of an impulse is a 2-D chirp), which makes the digests sensitive to phase errors anywhere."""
    rng = np.random.default_rng(seed + 1)
    r = rng.integers(0, x.shape[0], n_points)
    c = rng.integers(0, x.shape[1], n_points)
    x[r, c] += np.complex64(gain)
    return x


def large_csa_input(n_az: int, n_rg: int, seed: int) -> np.ndarray:
    return add_point_echoes(seeded_c64((n_az, n_rg), seed), seed)


def image_digest(img: np.ndarray, step: int, block: int = 128) -> dict:
    """What the large fixtures keep of a focused [n_rg, n_az] image: every ``step``-th sample of the flattened
    array, three whole rows and two whole columns, the energy of every ``block`` x ``block`` tile, and the
    whole-array sum / sum of squares."""
    n_rg, n_az = img.shape
    flat = img.reshape(-1)
    rows = (0, n_rg // 2 - 1, n_rg - 1)
    cols = (17 % n_az, (n_az // 2 + 4) % n_az)
    nb_r, nb_c = -(-n_rg // block), -(-n_az // block)
    e = np.zeros((nb_r, nb_c))
    for i in range(nb_r):
        blk = img[i * block:(i + 1) * block]
        p = (blk.real.astype(np.float64) ** 2 + blk.imag.astype(np.float64) ** 2)
        for j in range(nb_c):
            e[i, j] = p[:, j * block:(j + 1) * block].sum()
    return {"dec": np.ascontiguousarray(flat[::step]), "rows_idx": np.array(rows), "cols_idx": np.array(cols),
            "rows": np.stack([img[r] for r in rows]), "cols": np.stack([img[:, c] for c in cols]),
            "tile_energy": e, "sum": flat.sum(dtype=np.complex128), "sumsq": float(e.sum()),
            "shape": np.array(img.shape), "step": step, "block": block}
