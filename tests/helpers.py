"""Shared fixtures for the parity tests (inputs mirror the reference's tests)."""
from __future__ import annotations

import numpy as np

from oracle import analytic

UP = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]


def param_tree(kind="J2", hardening=("voce",), elastic=None, hill=None, a=None,
               active=("E", "nu", "D", "S", "Y"), rotation=None):
    """(values, active_flags, transforms=None-tree) for a J2AnalyticalProblem-style
    material (tests/support/test_problems.py:142-162) with optional variations."""
    values, act, _ = analytic.j2_voce_param_tree(kind)
    if elastic is not None:
        values["elastic"] = dict(elastic)
        act["elastic"] = {k: False for k in elastic}
    if hill is not None:
        values["plastic"]["effective stress"] = {"hill": dict(zip("FGHLMN", hill))}
        act["plastic"]["effective stress"] = {"hill": {k: False for k in "FGHLMN"}}
    if a is not None:
        values["plastic"]["effective stress"] = {"hosford": {"a": float(a)}}
    hd = {}
    ahd = {}
    if "voce" in hardening:
        hd["voce"] = {"S": 200.0, "D": 20.0}; ahd["voce"] = {"S": False, "D": False}
    if "linear" in hardening:
        hd["linear"] = {"K": 1500.0}; ahd["linear"] = {"K": False}
    values["plastic"]["flow stress"]["hardening"] = hd
    act["plastic"]["flow stress"]["hardening"] = ahd
    act["plastic"]["flow stress"]["initial yield"]["Y"] = False
    if rotation is not None:
        values["rotation matrix"] = np.asarray(rotation, float)
    # activate by short name
    def setact(tree, atree):
        for k in tree:
            if isinstance(tree[k], dict):
                setact(tree[k], atree[k])
            elif k in active and np.ndim(tree[k]) == 0:
                atree[k] = True
    setact(values, act)
    transforms = _const_like(values, None)
    return values, act, transforms


def _const_like(t, c):
    return {k: _const_like(v, c) for k, v in t.items()} if isinstance(t, dict) else c


def random_strains(rng, n, scale=1e-3, diag_only=False):
    d = rng.normal(size=(6, n))
    if diag_only:
        d[[1, 2, 4]] = 0.0
    nrm = np.sqrt(d[0] ** 2 + d[3] ** 2 + d[5] ** 2 + 2 * (d[1] ** 2 + d[2] ** 2 + d[4] ** 2))
    return d / nrm * rng.uniform(0.3, 5.0, size=n) * scale


def rotation_matrix(axis, angle):
    axis = np.asarray(axis, float); axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def rel_err(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
