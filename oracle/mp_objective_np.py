"""Batched CPU restatement of the material-point calibration objectives.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``cmad/objectives/mp_objective.py:62-147`` (forward pass + reverse-time adjoint)
and ``:158-215`` (forward sensitivities) with the ``Calibration`` QoI
(``cmad/qois/calibration.py:56-66``); every model derivative (dC/dxi,
dC/dxi_prev, dC/dp, dcauchy/dxi, dcauchy/dp) comes from the dual-number C++
oracle, i.e. is AD-derived like the reference's.  Vectorised over points with
NumPy; validated against the line-by-line torch oracle in tests/.
"""
from __future__ import annotations

import numpy as np

from . import oracle_c

_COMP = np.array([0, 1, 2, 1, 3, 4, 2, 4, 5])      # 3x3 row-major entry -> packed component


def forward(prob_solve, strain_hist, xi0):
    """newton_solve per step; strain_hist (N+1, 6|9, n); returns xi (N+1, 7, n), iters."""
    N = strain_hist.shape[0] - 1
    xi = np.zeros((N + 1,) + xi0.shape); xi[0] = xi0
    iters = np.zeros((N + 1, xi0.shape[1]), dtype=np.int32)
    for t in range(1, N + 1):
        r = oracle_c.mp_update(prob_solve, xi[t - 1], strain_hist[t], want=("xi", "iters"))
        xi[t], iters[t] = r["xi"], r["iters"]
    return xi, iters


def _step_terms(prob_eval, xi_t, xi_tm1, strain_t, data_t, weight):
    n = xi_t.shape[1]
    r = oracle_c.mp_update(prob_eval, xi_tm1, strain_t, xi_init=xi_t,
                           want=("sigma", "dC_dxi", "dC_dxi_prev", "dC_dp", "dsig_dxi", "dsig_dp"))
    na = len(prob_eval.active_pid)
    A = r["dC_dxi"].reshape(7, 7, n); B = r["dC_dxi_prev"].reshape(7, 7, n)
    dCdp = r["dC_dp"].reshape(7, na, n) if na else np.zeros((7, 0, n))
    sig9 = r["sigma"][_COMP]                                   # (9, n)
    mis = weight.reshape(9, 1) * (sig9 - data_t)               # calibration.py:64-66
    J = 0.5 * np.sum(mis * mis, axis=0)
    dJds9 = weight.reshape(9, 1) * mis                         # dJ/dsigma_ij
    ds9dx = r["dsig_dxi"].reshape(6, 7, n)[_COMP]              # (9, 7, n)
    dJdx = np.einsum("kn,kcn->cn", dJds9, ds9dx)
    if na:
        ds9dp = r["dsig_dp"].reshape(6, na, n)[_COMP]
        dJdp = np.einsum("kn,kcn->cn", dJds9, ds9dp)
    else:
        dJdp = np.zeros((0, n))
    return A, B, dCdp, J, dJdx, dJdp


def objective(values, active_idx, strain_hist, data_hist, weight, strategy="adjoint",
              newton=None):
    """(J, grad[native]) summed over points; also returns per-point J and grad."""
    newton = newton or dict(newton_mode="imperative", max_iters=10, abs_tol=1e-14, rel_tol=1e-14)
    sc = strain_hist.shape[1]
    prob_solve = oracle_c.describe(values, active_idx, strain_comps=sc, **newton)
    prob_eval = oracle_c.describe(values, active_idx, strain_comps=sc, newton_mode="imperative", max_iters=0)
    n = strain_hist.shape[2]; N = strain_hist.shape[0] - 1
    xi, iters = forward(prob_solve, strain_hist, np.zeros((7, n)))
    na = len(prob_eval.active_pid)
    Jp = np.zeros(n); g = np.zeros((na, n))
    weight = np.asarray(weight, float)
    solve = lambda M, b: np.linalg.solve(np.moveaxis(M, 2, 0), np.moveaxis(b, -1, 0)[..., None])[..., 0].T
    if strategy == "adjoint":
        hist = np.zeros((7, n))
        for t in range(N, 0, -1):
            A, B, dCdp, J, dJdx, dJdp = _step_terms(prob_eval, xi[t], xi[t - 1], strain_hist[t], data_hist[t], weight)
            Jp += J
            phi = solve(np.swapaxes(A, 0, 1), -dJdx + hist)          # mp_objective.py:129
            hist = -np.einsum("rcn,rn->cn", B, phi)                  # :134
            g += np.einsum("rn,rcn->cn", phi, dCdp) + dJdp           # :142
    else:
        X = np.zeros((7, na, n))
        for t in range(1, N + 1):
            A, B, dCdp, J, dJdx, dJdp = _step_terms(prob_eval, xi[t], xi[t - 1], strain_hist[t], data_hist[t], weight)
            Jp += J
            rhs = -dCdp - np.einsum("rqn,qcn->rcn", B, X)            # :205
            X = np.stack([solve(A, rhs[:, c]) for c in range(na)], axis=1) if na else X
            g += np.einsum("qn,qcn->cn", dJdx, X) + dJdp             # :208
    return float(Jp.sum()), g.sum(axis=1), Jp, g, xi, iters
