"""Materials / solver settings shared by the golden-vector generator
(`make_reference_golden.py`, which runs the reference's own code) and by the
tests that replay the stored inputs through the oracle and the CUDA path.
No dependency on the reference or on JAX."""
from __future__ import annotations

import numpy as np


def rotation_matrix(axis, angle):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


BARLAT_KEYS = tuple(f"{p}_{ij}" for p in ("sp", "dp") for ij in ("12", "13", "21", "23", "31", "32", "44", "55", "66")) + ("a",)


def material(kind):
    """Parameter `values` pytree in the reference's layout; E, nu, Y, S, D of
    tests/support/test_problems.py:151 unless noted."""
    el = {"E": 200e3, "nu": 0.3}
    Y, hd, Q = 200.0, {"voce": {"S": 200.0, "D": 20.0}}, np.eye(3)
    if kind == "J2":
        es = {"J2": 0.0}
    elif kind == "hill":                       # J2-equivalent Hill (test_problems.py:13)
        es = {"hill": {k: 0.5 for k in "FGHLMN"}}
    elif kind == "hill_rot":                   # anisotropic Hill, rotated material axes, voce + linear
        es = {"hill": dict(zip("FGHLMN", (0.45, 0.6, 0.55, 1.4, 1.6, 1.5)))}
        hd = {"voce": {"S": 200.0, "D": 20.0}, "linear": {"K": 1500.0}}
        Q = rotation_matrix([1.0, 2.0, -0.5], 0.7)
    elif kind == "hosford":                    # test_problems.py:15
        es = {"hosford": {"a": 4.0}}
    elif kind == "hosford_notch":              # examples/notch_hosford.yaml:37-42
        el = {"E": 1000.0, "nu": 0.25}
        es = {"hosford": {"a": 100.0}}
        Y, hd = 2.0, {"voce": {"S": 10.0, "D": 2.0}}
    elif kind in ("barlat", "barlat_rot", "barlat_a8"):
        # Yld2004-18p with the AL7079 fit of cmad/calibrations/al7079/support.py:80-89 (a = 18.2,
        # a non-integer exponent); "_a8": the same tensors with the fcc exponent 8; "_rot": rotated axes
        c = (0.4555, 1.0274, 0.7101, 1.3755, 0.5314, 0.8817, 1.0558, 1.1133, 0.9220,
             1.2431, 1.5438, 1.2204, 0.7632, 0.5327, 0.3015, 0.9722, 0.7399, 1.0760)
        es = {"barlat": dict(zip(BARLAT_KEYS, c + (8.0 if kind == "barlat_a8" else 18.2,)))}
        if kind == "barlat_rot":
            Q = rotation_matrix([0.3, -1.0, 0.8], 0.9)
            hd = {"voce": {"S": 200.0, "D": 20.0}, "linear": {"K": 1500.0}}
    elif kind == "J2_kappa_mu":                # another elastic-constant pair
        el = {"kappa": 200e3 / (3 * (1 - 0.6)), "mu": 200e3 / 2.6}
        es = {"J2": 0.0}
    else:
        raise ValueError(kind)
    return {"rotation matrix": Q, "elastic": el,
            "plastic": {"effective stress": es,
                        "flow stress": {"initial yield": {"Y": Y}, "hardening": hd}}}


def const_like(t, c):
    return {k: const_like(v, c) for k, v in t.items()} if isinstance(t, dict) else c


def active_all_scalars(values):
    """Every scalar leaf active except the J2 placeholder and the Hosford exponent."""
    act = const_like(values, False)

    def mark(t, a):
        for k in t:
            if isinstance(t[k], dict):
                mark(t[k], a[k])
            elif np.ndim(t[k]) == 0 and k not in ("J2", "a"):
                a[k] = True
    mark(values, act)
    return act


def active_kernel_set(values):
    """`active_all_scalars`, within the kernels' limit of CMADX_MAX_ACTIVE = 16 columns: a Barlat
    material (23 differentiable scalar leaves) keeps the elastic / flow-stress leaves and the nine
    `sp_*` coefficients (the `dp_*` columns and the exponent are covered by tests/test_barlat.py)."""
    act = active_all_scalars(values)
    b = act["plastic"]["effective stress"].get("barlat")
    if b is not None:
        for k in b:
            if k.startswith("dp_"):
                b[k] = False
    return act


def objective_trees(kind, scaled):
    """(values, active, transforms) of the objective fixtures: flow-stress
    parameters active with log / bounds transforms (test_problems.py:27-48), or
    elastic + flow-stress parameters active without transforms (configs[0])."""
    values = material(kind)
    act = const_like(values, False)
    tr = const_like(values, None)
    fs_act = act["plastic"]["flow stress"]
    fs_act["initial yield"]["Y"] = True
    for k in fs_act["hardening"]["voce"]:
        fs_act["hardening"]["voce"][k] = True
    if scaled:
        fs = tr["plastic"]["flow stress"]
        fs["initial yield"]["Y"] = np.array([values["plastic"]["flow stress"]["initial yield"]["Y"]])
        S, D = (values["plastic"]["flow stress"]["hardening"]["voce"][k] for k in "SD")
        fs["hardening"]["voce"]["S"] = np.array([0.5 * S, 1.5 * S])
        fs["hardening"]["voce"]["D"] = np.array([0.5 * D, 1.5 * D])
    else:
        act["elastic"] = {k: True for k in values["elastic"]}
        if kind.startswith("barlat"):          # a few coefficients of both tensors, and the exponent
            for k in ("sp_12", "sp_44", "dp_23", "dp_66", "a"):
                act["plastic"]["effective stress"]["barlat"][k] = True
    return values, act, tr


NEWTON = {
    "mp": dict(max_iters=10, abs_tol=1e-14, rel_tol=1e-14),                      # make_newton_solve defaults
    "fe": dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12),                      # global_residual.py:292-297
    "notch": dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12,                   # notch_hosford.yaml:30-35
                  line_search_settings={"max evals": 100}),
}
