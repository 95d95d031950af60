"""Synthetic strain histories of the benchmark configurations.

Per point i: two random unit directions d_i, d'_i in symmetric-strain space
(6 i.i.d. N(0,1) components xx,xy,xz,yy,yz,zz, normalised to unit Frobenius
norm with shear counted twice), amplitude a_i ~ U[0.5,5]*(Y/E); strain ramps
along d_i for the first ``leg`` steps and then along d'_i (non-proportional
second leg, as the two-leg histories of the reference's
tests/objectives/test_J2_fd_checks.py:266-289).  The generator is counter based
(splitmix64 of (seed, i, k) -> Box-Muller), so any sub-range of points can be
regenerated on the host for the CPU baseline without storing the whole set.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, i: np.ndarray, k: int) -> np.ndarray:
    """U(0,1) from the counter (seed, i, k); never exactly 0."""
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(k))
        z = splitmix64(i.astype(np.uint64) * np.uint64(0xD1342543DE82EF95) + key)
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normals(seed: int, i: np.ndarray, k0: int, count: int) -> np.ndarray:
    out = np.empty((count, i.size))
    for c in range(0, count, 2):
        u1, u2 = uniform01(seed, i, k0 + c), uniform01(seed, i, k0 + c + 1)
        r = np.sqrt(-2.0 * np.log(u1))
        out[c] = r * np.cos(2.0 * np.pi * u2)
        if c + 1 < count:
            out[c + 1] = r * np.sin(2.0 * np.pi * u2)
    return out


def _unit(d: np.ndarray, diag_only: bool) -> np.ndarray:
    if diag_only:
        d[[1, 2, 4]] = 0.0
    nrm = np.sqrt(d[0] ** 2 + d[3] ** 2 + d[5] ** 2 + 2.0 * (d[1] ** 2 + d[2] ** 2 + d[4] ** 2))
    return d / nrm


def path_params(seed: int, i0: int, n: int, yield_strain: float = 1e-3,
                diag_only: bool = False, stride: int = 1):
    """(d (6,n), d2 (6,n), a (n)) for points i0, i0 + stride, ... (n of them)."""
    i = np.uint64(i0) + np.arange(n, dtype=np.uint64) * np.uint64(stride)
    d = _unit(normals(seed, i, 0, 6), diag_only)
    d2 = _unit(normals(seed, i, 8, 6), diag_only)
    a = (0.5 + 4.5 * uniform01(seed, i, 16)) * yield_strain
    return d, d2, a


def strain_at_step(d, d2, a, t: int, leg: int = 50):
    """Works on NumPy arrays and torch tensors alike (same IEEE op order)."""
    if t <= leg:
        return (a * (t / leg)) * d
    return (a * 1.0) * d + (a * ((t - leg) / leg)) * d2
