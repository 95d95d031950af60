"""Closed-form J2 + Voce proportional-loading path: the reference's own
known-answer generator (KA1), restated in NumPy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
Follows ``cmad/verification/solutions.py:30-58`` (``compute_plastic_fields``)
and ``cmad/verification/functions.py:7-22`` (``J2_yield`` / ``J2_yield_normal``);
fixture values from ``tests/support/test_problems.py:142-162`` and
``tests/models/test_elastic_plastic_models.py:15-61``.
"""
from __future__ import annotations

import numpy as np

# E, nu, Y, S, D  (tests/support/test_problems.py:151)
J2_ANALYTICAL_PARAMS = np.array([200e3, 0.3, 200., 200., 20.])


def j2_yield(cauchy: np.ndarray) -> float:
    s = cauchy - np.trace(cauchy) / 3. * np.eye(3)
    return float(np.sqrt(1.5) * np.linalg.norm(s))


def j2_yield_normal(cauchy: np.ndarray) -> np.ndarray:
    s = cauchy - np.trace(cauchy) / 3. * np.eye(3)
    return np.sqrt(1.5) * s / np.linalg.norm(s)


def plastic_fields(stress_mask: np.ndarray, isotropic_params=J2_ANALYTICAL_PARAMS,
                   max_alpha: float = 0.5, num_steps: int = 100):
    """Stress, total strain and alpha histories (each with ``num_steps``
    samples) of a point loaded proportionally along ``stress_mask`` with alpha
    prescribed on a uniform grid: sigma = mask*(Y + S(1-exp(-D alpha)))/phi(mask),
    plastic strain integrated by backward Euler along the (constant) normal,
    total strain = Hooke^-1(sigma) + plastic strain."""
    E, nu, Y, S, D = isotropic_params
    alpha = np.linspace(0., max_alpha, num_steps)
    dalpha = alpha[1] - alpha[0]
    level = (Y + S * (1. - np.exp(-D * alpha))) / j2_yield(stress_mask)
    stress = stress_mask[:, :, None] * level[None, None, :]
    pstrain = np.zeros((3, 3, num_steps))
    for k in range(1, num_steps):
        pstrain[:, :, k] = pstrain[:, :, k - 1] + dalpha * j2_yield_normal(stress[:, :, k])
    tr = np.einsum("iik->k", stress)
    strain = (stress - nu * (tr[None, None, :] * np.eye(3)[:, :, None] - stress)) / E + pstrain
    return stress, strain, alpha


def stress_masks_3d():
    """tests/models/test_elastic_plastic_models.py:45-55: uniaxial and
    equal-and-opposite biaxial."""
    m0 = np.zeros((3, 3)); m0[0, 0] = 1.
    m1 = np.eye(3); m1[1, 1] = -1.; m1[2, 2] = 0.
    return [m0, m1]


def deformation_gradient_history(strain: np.ndarray, ndims: int = 3) -> np.ndarray:
    """``get_F`` (test_elastic_plastic_models.py:37-42): F[...,0]=I, F[...,k+1]=I+strain[...,k]."""
    n = strain.shape[2]
    F = np.repeat(np.eye(ndims)[:, :, None], n + 1, axis=2)
    F[:, :, 1:] += strain[:ndims, :ndims, :]
    return F


def j2_voce_param_tree(effective_stress: str = "J2", flat=J2_ANALYTICAL_PARAMS):
    """Parameter pytrees of ``params_J2_voce`` (tests/support/test_problems.py:9-113)
    with ``scale_params=True`` transforms: returns (values, active_flags, transforms)."""
    E, nu, Y, S, D = [float(v) for v in flat]
    if effective_stress == "J2":
        es = {"J2": 0.}
    elif effective_stress == "hill":
        es = {"hill": {k: 0.5 for k in "FGHLMN"}}
    elif effective_stress == "hosford":
        es = {"hosford": {"a": 4.}}
    else:
        raise ValueError(effective_stress)
    values = {
        "rotation matrix": np.eye(3),
        "elastic": {"E": E, "nu": nu},
        "plastic": {
            "effective stress": es,
            "flow stress": {"initial yield": {"Y": Y},
                            "hardening": {"voce": {"S": S, "D": D}}}}}

    def const_like(t, c):
        return {k: const_like(v, c) for k, v in t.items()} if isinstance(t, dict) else c

    active = const_like(values, False)
    active["plastic"]["flow stress"] = const_like(values["plastic"]["flow stress"], True)
    transforms = const_like(values, None)
    fs = transforms["plastic"]["flow stress"]
    fs["initial yield"]["Y"] = np.array([200.])
    fs["hardening"]["voce"]["S"] = np.array([100., 300.])
    fs["hardening"]["voce"]["D"] = np.array([10., 30.])
    return values, active, transforms
