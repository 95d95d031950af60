"""FE quantities of interest accumulated over the quasi-static time loop, with the pieces the
gradient drivers need (cmad/qois/fe_qoi.py: ``J = sum_n J_n(U_n, U_{n-1}, xi_n, xi_{n-1}, t_n,
t_{n-1})``).  Host-side, NumPy: the QoIs read the solved state; the constitutive work stays in
the element kernels the drivers call.

  FEDisplacementL2      cmad/qois/fe_displacement_l2.py:106-123
  FEDisplacementMatch   cmad/qois/fe_displacement_match.py:21-150
  FELoadMatch           cmad/qois/fe_load_match.py:24-196 (match mode; `reaction_series` = write mode)
  FEWeightedSum         cmad/qois/fe_weighted_sum.py:21-78

A QoI exposes ``value(s)``, ``dU(s)`` (explicit dJ_n/dU_n) and ``dR(s)`` (dJ_n/dR, for QoIs that
read reactions: the assembled residual depends on the parameters directly and on xi_{n-1}, which
the adjoint / direct drivers propagate through the K6 kernels).  ``s`` is a :class:`StepState`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np


@dataclass
class StepState:
    U: np.ndarray
    t: float
    t_prev: float
    R: np.ndarray | None = None          # assembled (un-embedded) residual at the converged state
    step: int = 0                        # index of t in the time schedule


class FEQoI:
    needs_residual = False

    def value(self, s: StepState) -> float:
        raise NotImplementedError

    def dU(self, s: StepState) -> np.ndarray:
        raise NotImplementedError

    def dR(self, s: StepState) -> np.ndarray | None:
        return None


def _ip_interp(N, elem_eq, U):
    U_e = np.asarray(U)[elem_eq].reshape(elem_eq.shape[0], -1, 3)
    return np.einsum("pa,eak->epk", N, U_e)


class FEDisplacementL2(FEQoI):
    """``J = 1/(T |Omega|) sum_n dt_n int |u_n|^2 dV``."""

    def __init__(self, arrays, t_schedule: Sequence[float], weight: float = 1.0):
        self.N, self.eq = arrays.N.cpu().numpy(), arrays.elem_eq.cpu().numpy().astype(np.int64)
        self.wdet = (arrays.det * arrays.quad_w[None, :]).cpu().numpy()
        self.c = float(weight) / ((float(t_schedule[-1]) - float(t_schedule[0])) * self.wdet.sum())
        self.n_dofs = arrays.n_dofs

    def _diff(self, s):
        return _ip_interp(self.N, self.eq, s.U)

    def value(self, s):
        d = self._diff(s)
        return self.c * (s.t - s.t_prev) * float(((d * d).sum(axis=-1) * self.wdet).sum())

    def dU(self, s):
        g_e = 2.0 * np.einsum("pa,epk,ep->eak", self.N, self._diff(s), self.wdet)
        g = np.zeros(self.n_dofs)
        np.add.at(g, self.eq.reshape(-1), g_e.reshape(-1))
        return self.c * (s.t - s.t_prev) * g


class FEDisplacementMatch(FEDisplacementL2):
    """``J = w/(T |Omega|) sum_n dt_n int |u_n - u_data_n|^2 dV``; ``data (num_steps, n_nodes, 3)``
    holds one nodal displacement field per schedule time (the initial time included)."""

    def __init__(self, arrays, t_schedule: Sequence[float], data, weight: float = 1.0):
        super().__init__(arrays, t_schedule, weight)
        data = np.asarray(data, dtype=np.float64)
        if data.shape[0] != len(t_schedule):
            raise ValueError(f"FEDisplacementMatch: data has {data.shape[0]} steps but the time schedule has "
                             f"{len(t_schedule)} (expected one displacement field per schedule time, including "
                             f"the initial time)")
        self.data = data.reshape(len(t_schedule), -1)
        if self.data.shape[1] != self.n_dofs:
            raise ValueError(f"FEDisplacementMatch: data flattens to {self.data.shape[1]} dofs/step but the "
                             f"problem has {self.n_dofs} total dofs")

    def _diff(self, s):
        return _ip_interp(self.N, self.eq, s.U - self.data[s.step])


class FELoadMatch(FEQoI):
    """``J = w/T sum_n dt_n sum_c (R_{c,n} - d_{c,n})^2`` with ``R_c`` the assembled residual summed
    over the Dirichlet-prescribed dofs of component ``c`` on a side set (``eq_per_component``)."""
    needs_residual = True

    def __init__(self, eq_per_component: Sequence[np.ndarray], t_schedule: Sequence[float], data, weight: float = 1.0):
        self.eqs = [np.asarray(e, dtype=np.int64) for e in eq_per_component]
        data = np.asarray(data, dtype=np.float64)
        if data.ndim == 1 and len(self.eqs) == 1:
            data = data.reshape(-1, 1)
        if data.shape != (len(t_schedule), len(self.eqs)):
            raise ValueError(f"FELoadMatch: data has shape {data.shape} but expected (num_steps={len(t_schedule)}, "
                             f"num_components={len(self.eqs)})")
        self.data = data
        self.c = float(weight) / (float(t_schedule[-1]) - float(t_schedule[0]))

    def reaction(self, R) -> np.ndarray:
        return np.array([np.asarray(R)[e].sum() for e in self.eqs])

    def value(self, s):
        m = self.reaction(s.R) - self.data[s.step]
        return self.c * (s.t - s.t_prev) * float(m @ m)

    def dU(self, s):
        return np.zeros_like(s.U)

    def dR(self, s):
        m = self.reaction(s.R) - self.data[s.step]
        g = np.zeros_like(np.asarray(s.R, dtype=np.float64))
        for e, mc in zip(self.eqs, m):
            g[e] += 2.0 * self.c * (s.t - s.t_prev) * mc
        return g


class FEWeightedSum(FEQoI):
    """Sum of sub-QoIs, each carrying its own weight."""

    def __init__(self, terms: Sequence[FEQoI]):
        self.terms = list(terms)
        self.needs_residual = any(t.needs_residual for t in self.terms)

    def value(self, s):
        return float(sum(t.value(s) for t in self.terms))

    def dU(self, s):
        return sum(t.dU(s) for t in self.terms)

    def dR(self, s):
        parts = [g for g in (t.dR(s) for t in self.terms) if g is not None]
        return sum(parts) if parts else None


def reaction_series(qoi: FELoadMatch, residuals: Sequence[Any]) -> np.ndarray:
    """The write mode of ``fe_load_match`` (cmad/qois/fe_load_match.py:146-177): the reaction per
    component at every stored step, from the assembled residuals of those steps."""
    return np.array([qoi.reaction(R) for R in residuals])
