"""Quasi-static FE driver on top of the element-block kernels: global Newton with the
embedded Dirichlet treatment and cubic backtracking line search, and the load-step
loop with a step-QoI - the direct callers of the hot path.

Host-side Python like the reference's; mirrors, with the same settings keys and defaults:

  ``fe_newton_solve``        cmad/fem/nonlinear_solver.py:188-293 (``_fe_newton_primal``)
  embedded BCs               cmad/fem/sparse_solve.py:1058-1174 (``_embedded_bc_enforce``,
                             ``_embedded_residual``)
  line search (cubic model)  cmad/util/line_search.py:49-71, 95-189
  ``fe_quasistatic_drive``   cmad/fem/driver.py:103-146 (scan over load steps)
  ``FEDisplacementL2``       cmad/qois/fe_displacement_l2.py:78-125

Every assembly (the Newton iterates AND the line-search probes) is one K3 + K5 call on
the device through ``assemble``; the sparse linear solve is the reference's default
``direct`` solver - SciPy SuperLU on the host (cmad/fem/sparse_solve.py:89, reached
there through ``jax.pure_callback``) - and, like every global sparse solver of the
reference, is outside the B200 path.  The driver is written against an
``assemble(U, xi_prev) -> (R, K_data, xi)`` callable so that the tests can run the very
same loop over the CPU oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

DEFAULT_LINE_SEARCH = {"max evals": 4, "sufficient decrease": 1.0e-4,
                       "min backtrack factor": 0.5, "max backtrack factor": 0.9}
# _FE_RESIDUALS_DEFAULTS["global residual"], cmad/io/deck.py:57-70
DEFAULT_NONLINEAR = {"max iters": 10, "abs tol": 1.0e-12, "rel tol": 1.0e-12,
                     "line search": DEFAULT_LINE_SEARCH}


def cubic_min(phi_0, dphi_0, a, phi, slope_a):
    """line_search.py:49-71: minimiser of the two-point Hermite cubic."""
    d1 = dphi_0 + slope_a - 3.0 * (phi_0 - phi) / (0.0 - a)
    radicand = d1 * d1 - dphi_0 * slope_a
    d2 = math.sqrt(max(radicand, 0.0))
    denom = slope_a - dphi_0 + 2.0 * d2
    if radicand < 0.0 or denom == 0.0:
        return 0.5 * a
    return a - a * (slope_a + d2 - d1) / denom


def line_search(eval_fn, phi_0, dphi_0, settings, init_aux):
    """line_search.py:95-189 with trial slopes (cubic contraction).
    ``eval_fn(alpha) -> (phi, slope, aux)``; returns ``(alpha, aux, n_evals)``."""
    s = {**DEFAULT_LINE_SEARCH, **(settings or {})}
    armijo = s["sufficient decrease"] * dphi_0
    n, alpha, accepted, aux = 0, 1.0, False, init_aux
    best_alpha, best_phi, best_aux = 1.0, math.inf, init_aux
    while n < s["max evals"] and not accepted:
        phi, slope, aux = eval_fn(alpha)
        finite = math.isfinite(phi)
        if finite and phi < best_phi:
            best_alpha, best_phi, best_aux = alpha, phi, aux
        accepted = finite and phi <= phi_0 + alpha * armijo
        model = cubic_min(phi_0, dphi_0, alpha, phi, slope) if finite else float("nan")
        lo, hi = s["min backtrack factor"] * alpha, s["max backtrack factor"] * alpha
        contracted = model if model != model else min(max(model, lo), hi)
        if not accepted:
            alpha = contracted if finite else 0.5 * alpha
        n += 1
    return (alpha, aux, n) if accepted else (best_alpha, best_aux, n)


@dataclass
class DirichletBCs:
    """Prescribed dofs and their values as a function of pseudo-time (the deck's
    ``dirichlet bcs`` expressions evaluated per load step, cmad/fem/dof.py)."""
    indices: np.ndarray                              # (n_presc,) global equations
    values: Callable[[float], np.ndarray]            # t -> (n_presc,)


@dataclass
class SparsePattern:
    """Deduplicated COO pattern of the assembled tangent (``coo_rows`` / ``coo_cols``,
    cmad/fem/kernel_arrays.py:81-90)."""
    rows: np.ndarray
    cols: np.ndarray
    n: int

    def csr(self, data: np.ndarray) -> sp.csr_matrix:
        return sp.csr_matrix((data, (self.rows, self.cols)), shape=(self.n, self.n))


def embedded_system(pattern: SparsePattern, K_data, R, U, bcs: DirichletBCs, t: float):
    """``(r, K_emb)`` of sparse_solve.py:1058-1174: prescribed rows and columns of K
    zeroed, the assembled diagonal kept at the prescribed rows; the (free, prescribed)
    coupling moved to the right-hand side; prescribed rows of r = K_ii (U - value)."""
    K = pattern.csr(np.asarray(K_data))
    idx = bcs.indices
    vals = np.asarray(bcs.values(t), dtype=np.float64)
    K_ii = K.diagonal()[idx]
    inc = np.zeros(pattern.n)
    inc[idx] = vals - U[idx]
    r = np.asarray(R) + K @ inc
    r[idx] = K_ii * (U[idx] - vals)
    keep = np.ones(pattern.n)
    keep[idx] = 0.0
    D = sp.diags(keep)
    K_emb = (D @ K @ D + sp.csr_matrix((K_ii, (idx, idx)), shape=K.shape)).tocsc()
    return r, K_emb


@dataclass
class NewtonLog:
    iters: int = 0
    assemblies: int = 0
    K_emb: object = None          # embedded tangent at the returned state (for sensitivities)
    residual_norms: list = field(default_factory=list)
    alphas: list = field(default_factory=list)


def fe_newton_solve(assemble, pattern: SparsePattern, bcs: DirichletBCs, U_prev: np.ndarray,
                    xi_prev, t: float, settings: dict | None = None):
    """One load step: ``(U*, xi*, log)``.  ``assemble(U, xi_prev) -> (R, K_data, xi)``
    with ``R (n_dofs,)`` and ``K_data`` on ``pattern`` as host arrays (``xi`` is opaque
    to the driver and stays wherever ``assemble`` keeps it).  An assembler with an
    ``enforced(U, xi_prev, t) -> (r, K_emb, xi)`` attribute (:class:`DeviceEmbeddedBCs`) applies
    the embedded-BC treatment itself, on the device."""
    s = {**DEFAULT_NONLINEAR, **(settings or {})}
    ls = {**DEFAULT_LINE_SEARCH, **(s.get("line search") or {})}
    log = NewtonLog()
    enforced = getattr(assemble, "enforced", None)

    def assemble_enforced(U):
        log.assemblies += 1
        if enforced is not None:
            return enforced(U, xi_prev, t)
        R, K_data, xi = assemble(U, xi_prev)
        r, K_emb = embedded_system(pattern, K_data, R, U, bcs, t)
        return r, K_emb, xi

    U = np.array(U_prev, dtype=np.float64)
    r, K, xi = assemble_enforced(U)
    R0 = max(float(np.linalg.norm(r)), s["abs tol"])
    i = 0
    while True:
        nrm = float(np.linalg.norm(r))
        log.residual_norms.append(nrm)
        if not (i < s["max iters"] and nrm >= s["abs tol"] and nrm >= s["rel tol"] * R0):
            break
        dU = spla.splu(K).solve(-r)
        if ls["max evals"] > 0:
            rr = float(r @ r)

            def eval_fn(alpha, U=U, dU=dU):
                rt, Kt, xit = assemble_enforced(U + alpha * dU)
                return 0.5 * float(rt @ rt), float(rt @ (Kt @ dU)), (rt, Kt, xit)

            alpha, (r, K, xi), _ = line_search(eval_fn, 0.5 * rr, -rr, ls, (r, K, xi))
            U = U + alpha * dU
            log.alphas.append(alpha)
        else:
            U = U + dU
            r, K, xi = assemble_enforced(U)
        i += 1
    log.iters = i
    log.K_emb = K
    return U, xi, log


def displacement_l2_step(N: np.ndarray, wdet: np.ndarray, elem_eq: np.ndarray, U: np.ndarray) -> float:
    """``sum_e sum_ip |N u_e|^2 w det`` (fe_displacement_l2.py:106-123)."""
    U_e = np.asarray(U)[elem_eq].reshape(elem_eq.shape[0], -1, 3)
    u_ip = np.einsum("pa,eak->epk", N, U_e)
    return float(((u_ip * u_ip).sum(axis=-1) * wdet).sum())


def _step_state(qoi, assemble, U, xi_prev, t, t_prev, k):
    """StepState of a converged load step; QoIs that read reactions get the assembled (un-embedded)
    residual and tangent at the converged state from one more assembly (as the reference's
    FELoadMatch._reaction_at re-runs the residual assembly, cmad/qois/fe_load_match.py:179-196)."""
    from .fe_qoi import StepState
    s = StepState(U=U, t=t, t_prev=t_prev, step=k)
    K_data = None
    if qoi is not None and qoi.needs_residual:
        R, K_data, _ = assemble(U, xi_prev)
        s.R = np.asarray(R, dtype=np.float64)
    return s, K_data


def fe_quasistatic_drive(assemble, pattern: SparsePattern, bcs: DirichletBCs, U0: np.ndarray, xi0,
                         t_schedule: Sequence[float], settings: dict | None = None,
                         step_qoi: Callable[[np.ndarray, float, float], float] | None = None, qoi=None):
    """Load-step loop (driver.py:103-146): returns ``(U_steps, xi_last, J, logs)``;
    ``step_qoi(U, t, t_prev)`` - or a :class:`cmad_b200.fe_qoi.FEQoI` object ``qoi`` - is summed into
    ``J`` after every converged step."""
    U, xi = np.array(U0, dtype=np.float64), xi0
    U_steps, logs, J = [], [], 0.0
    for k in range(1, len(t_schedule)):
        t, t_prev = float(t_schedule[k]), float(t_schedule[k - 1])
        xi_prev = xi
        U, xi, log = fe_newton_solve(assemble, pattern, bcs, U, xi_prev, t, settings)
        U_steps.append(U.copy())
        logs.append(log)
        if qoi is not None:
            J += qoi.value(_step_state(qoi, assemble, U, xi_prev, t, t_prev, k)[0])
        elif step_qoi is not None:
            J += step_qoi(U, t, t_prev)
    return np.array(U_steps), xi, J, logs


def displacement_l2_step_dU(N: np.ndarray, wdet: np.ndarray, elem_eq: np.ndarray, U: np.ndarray) -> np.ndarray:
    """Gradient of :func:`displacement_l2_step` w.r.t. the global ``U``."""
    U_e = np.asarray(U)[elem_eq].reshape(elem_eq.shape[0], -1, 3)
    u_ip = np.einsum("pa,eak->epk", N, U_e)
    g_e = 2.0 * np.einsum("pa,epk,ep->eak", N, u_ip, wdet)
    g = np.zeros(np.asarray(U).shape[0])
    np.add.at(g, elem_eq.reshape(-1), g_e.reshape(-1))
    return g


def fe_direct_gradient(assemble, jvp, pattern: SparsePattern, bcs: DirichletBCs, U0, xi0, dxi0,
                       t_schedule: Sequence[float], n_active: int, settings: dict | None,
                       step_qoi=None, step_qoi_dU=None, qoi=None):
    """``(J, dJ/dp)`` in native parameter values by forward (direct) sensitivities through
    the load steps - the discrete equivalent of differentiating the trajectory of
    cmad/fem/driver.py:103-146 through the IFT rule of the FE Newton
    (cmad/fem/nonlinear_solver.py:450-542): per step and active parameter c,
        dR_c = JVP(e_c, dxi_prev_c) at fixed U*          (K6, device)
        K_emb dU_c = -dR_c   on the free dofs            (host SuperLU, factored once per step)
        dxi_c = JVP(e_c, dxi_prev_c, dU_c).xi            (K6 with the displacement direction)
        dJ_c += dq/dU . dU_c
    ``jvp(U, xi_prev, xi_state, c, dxi_prev_c, dU) -> (dR (n_dofs,) host, dxi)``;
    ``dxi0(c)`` gives the initial (zero) state sensitivities in the assembler's format."""
    U, xi = np.array(U0, dtype=np.float64), xi0
    dX = [dxi0(c) for c in range(n_active)]
    J, grad = 0.0, np.zeros(n_active)
    for k in range(1, len(t_schedule)):
        t, t_prev = float(t_schedule[k]), float(t_schedule[k - 1])
        xi_prev = xi
        U, xi, log = fe_newton_solve(assemble, pattern, bcs, U, xi_prev, t, settings)
        lu = spla.splu(log.K_emb)
        sbar = None
        if qoi is not None:
            st, _ = _step_state(qoi, assemble, U, xi_prev, t, t_prev, k)
            dq, sbar = qoi.dU(st), qoi.dR(st)
            J += qoi.value(st)
        else:
            dq = step_qoi_dU(U, t, t_prev)
            J += step_qoi(U, t, t_prev)
        for c in range(n_active):
            dR, _ = jvp(U, xi_prev, xi, c, dX[c], None)
            rhs = -np.asarray(dR)
            rhs[bcs.indices] = 0.0                     # prescribed values do not depend on p
            dU = lu.solve(rhs)
            dR_total, dX[c] = jvp(U, xi_prev, xi, c, dX[c], dU)      # dR/dp + dR/dxi_prev dxi_prev + K dU
            grad[c] += float(dq @ dU)
            if sbar is not None:
                grad[c] += float(sbar @ np.asarray(dR_total))
    return J, grad


def fe_adjoint_gradient(assemble, vjp, vjp_disp, pattern: SparsePattern, bcs: DirichletBCs, U0, xi0,
                        t_schedule: Sequence[float], n_active: int, settings: dict | None,
                        step_qoi=None, step_qoi_dU=None, qoi=None):
    """``(J, dJ/dp)`` in native parameter values by the DISCRETE ADJOINT through the load
    steps - what ``jax.grad`` of the trajectory of cmad/fem/driver.py:103-146 computes through
    the IFT rules of the FE Newton (cmad/fem/nonlinear_solver.py:450-542) and of the local
    Newton (cmad/models/nonlinear_solver.py:158-171).  Forward: the load steps, keeping
    ``(U_k, xi_k, K_emb_k)``.  Backward, k = N..1, with ``xbar`` the cotangent of ``xi_k``:
        ub        = (d xi_k/dU_k)^T xbar                     (K6 ``vjp_disp``, device)
        K_emb^T lam = -(dq_k/dU + ub)  on the free dofs      (host SuperLU)
        pbar, xbar <- VJP(Rbar = lam, xibar = xbar)          (K6 ``vjp``, device)
        grad += pbar
    One sparse solve per step instead of one per parameter (:func:`fe_direct_gradient`).
    ``vjp(U, xi_prev, xi_state, Rbar, xibar) -> (pbar (n_active,), xibar_prev)``;
    ``vjp_disp(U, xi_prev, xi_state, xibar) -> ubar (n_dofs,)``; ``xibar`` may be None (zero)."""
    U, xi = np.array(U0, dtype=np.float64), xi0
    J = 0.0
    steps = []
    for k in range(1, len(t_schedule)):
        t, t_prev = float(t_schedule[k]), float(t_schedule[k - 1])
        xi_prev = xi
        U, xi, log = fe_newton_solve(assemble, pattern, bcs, U, xi_prev, t, settings)
        sbar, Kt_sbar = None, None
        if qoi is not None:
            st, K_data = _step_state(qoi, assemble, U, xi_prev, t, t_prev, k)
            J += qoi.value(st)
            dq, sbar = qoi.dU(st), qoi.dR(st)
            if sbar is not None:                        # a QoI of the reactions: dJ_n/dU gets K^T sbar
                Kt_sbar = pattern.csr(np.asarray(K_data)).T @ sbar
        else:
            J += step_qoi(U, t, t_prev)
            dq = step_qoi_dU(U, t, t_prev)
        steps.append((U.copy(), xi_prev, xi, log.K_emb, dq, sbar, Kt_sbar))
    grad = np.zeros(n_active)
    xbar = None
    for U, xi_prev, xi, K_emb, dq, sbar, Kt_sbar in reversed(steps):
        rhs = np.asarray(dq, dtype=np.float64).copy()
        if Kt_sbar is not None:
            rhs += Kt_sbar
        if xbar is not None:
            rhs += np.asarray(vjp_disp(U, xi_prev, xi, xbar))
        rhs[bcs.indices] = 0.0                          # prescribed values do not depend on p
        lam = spla.splu(sp.csc_matrix(K_emb.T)).solve(-rhs)
        lam[bcs.indices] = 0.0
        # the residual's direct dependence on (p, xi_prev) is weighted by lam (equilibrium) and by
        # sbar (the QoI reads reactions): one VJP with Rbar = lam + sbar
        pbar, xbar = vjp(U, xi_prev, xi, lam if sbar is None else lam + sbar, xbar)
        grad += np.asarray(pbar)
    return J, grad


def cuda_vjp(material, arrays, active_pid, stab_mult: float | None = None):
    """``(vjp, vjp_disp)`` callables of :func:`fe_adjoint_gradient` over the K6 reverse-mode
    kernels (``stab_mult``: the mixed u-p formulation)."""
    import torch
    from . import fe
    dev = arrays.grad_N.device
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    u_plan = fe.disp_cotangent_plan(arrays, device=dev)

    def vjp(U, xi_prev, xi_state, Rbar, xibar):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        Rd = torch.from_numpy(np.ascontiguousarray(Rbar)).to(dev)
        pbar, xbp = fe.fe_block_vjp(material, arrays, Ud, xi_prev, xi_state, pid, Rd, xibar, stab_mult=stab_mult)
        return pbar.cpu().numpy(), xbp

    def vjp_disp(U, xi_prev, xi_state, xibar):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        ub = fe.fe_block_vjp_disp(material, arrays, Ud, xi_prev, xi_state, xibar, stab_mult=stab_mult)
        return u_plan.sum(ub.reshape(-1)).cpu().numpy()

    return vjp, vjp_disp


def cuda_jvp(material, arrays, r_plan, active_pid):
    """``jvp`` callable of :func:`fe_direct_gradient` over the K6 kernels."""
    import torch
    from . import fe
    dev = arrays.grad_N.device
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)

    def jvp(U, xi_prev, xi_state, c, dxi_prev, dU):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        dp = np.zeros(len(pid)); dp[c] = 1.0
        dUd = torch.from_numpy(np.ascontiguousarray(dU)).to(dev) if dU is not None else None
        o = fe.fe_block_jvp(material, arrays, Ud, xi_prev, xi_state, pid, dp, dxi_prev, dU=dUd)
        return r_plan.sum(o["R_elem"].reshape(-1)).cpu().numpy(), o["xi"]

    return jvp


def cuda_jvp_mixed(material, arrays, r_plan, active_pid, stab_mult: float = 1.0):
    """``jvp`` callable of :func:`fe_direct_gradient` for the mixed u-p formulation (K6 over
    both residual blocks); ``r_plan`` = :func:`cmad_b200.fe.mixed_r_plan`."""
    import torch
    from . import fe
    dev = arrays.grad_N.device
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    n_e, nu, npd = arrays.n_elems, 3 * arrays.n_basis, arrays.n_basis

    def jvp(U, xi_prev, xi_state, c, dxi_prev, dU):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        dp = np.zeros(len(pid)); dp[c] = 1.0
        dUd = torch.from_numpy(np.ascontiguousarray(dU)).to(dev) if dU is not None else None
        R_elem = torch.empty(n_e * (nu + npd), dtype=torch.float64, device=dev)          # [dR_u | dR_p]
        o = fe.fe_block_jvp(material, arrays, Ud, xi_prev, xi_state, pid, dp, dxi_prev, dU=dUd,
                            stab_mult=stab_mult,
                            out={"R_elem": R_elem[:n_e * nu].view(n_e, nu), "R_p_elem": R_elem[n_e * nu:].view(n_e, npd)})
        return r_plan.sum(R_elem).cpu().numpy(), o["xi"]

    return jvp


def cuda_assembler(material, newton, arrays, r_plan, k_plan, outputs=None):
    """``assemble`` callable over the CUDA kernels for one element block: K3 (R_e, K_e,
    xi) + K5 (deterministic R scatter, COO dedup); ``xi`` stays on the device."""
    import torch
    from . import fe
    dev = arrays.grad_N.device

    def device(U, xi_prev):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        R, vals, xi = fe.assemble_element_block(material, newton, arrays, Ud, xi_prev, r_plan=r_plan)
        K_data = k_plan.sum(vals)
        if outputs is not None:
            outputs["last"] = (R, K_data, xi)
        return R, K_data, xi

    def assemble(U, xi_prev):
        R, K_data, xi = device(U, xi_prev)
        return R.cpu().numpy(), K_data.cpu().numpy(), xi

    assemble.device = device
    return assemble


def cuda_assembler_mixed(material, newton, arrays, r_plan, k_plan, stab_mult: float = 1.0, outputs=None):
    """``assemble`` callable for the mixed u-p formulation (K3 with the momentum stress
    dev(cauchy) - p I + the pressure-block kernel + K5); ``U`` holds the block-major (u, p)
    dofs, ``k_plan`` the dedup scatter of :func:`cmad_b200.fe_mesh.coo_pattern_mixed`."""
    import torch
    from . import fe
    dev = arrays.grad_N.device

    def device(U, xi_prev):
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(dev)
        R, vals, xi = fe.assemble_element_block_mixed(material, newton, arrays, Ud, xi_prev,
                                                      stab_mult=stab_mult, r_plan=r_plan)
        K_data = k_plan.sum(vals)
        if outputs is not None:
            outputs["last"] = (R, K_data, xi)
        return R, K_data, xi

    def assemble(U, xi_prev):
        R, K_data, xi = device(U, xi_prev)
        return R.cpu().numpy(), K_data.cpu().numpy(), xi

    assemble.device = device
    return assemble


class DeviceEmbeddedBCs:
    """Wraps a CUDA assembler so that the embedded-BC treatment (sparse_solve.py:1058-1174)
    also runs on the device (``fe.EmbeddedBCPlan``): per assembly, K3 + K5 + two HBM-bound
    passes, then ONE device->host copy of ``(r, K_emb data)`` for the host sparse solve.
    The CSC structure of ``K_emb`` is fixed by the pattern and built once."""

    def __init__(self, assemble, pattern: SparsePattern, bcs: DirichletBCs, device):
        import torch
        from . import fe
        self._assemble, self._bcs, self._dev = assemble, bcs, torch.device(device)
        self._plan = fe.EmbeddedBCPlan(pattern.rows, pattern.cols, pattern.n, bcs.indices, device=device)
        ids = sp.csc_matrix((np.arange(1, len(pattern.rows) + 1, dtype=np.float64), (pattern.rows, pattern.cols)),
                            shape=(pattern.n, pattern.n))
        ids.sort_indices()
        self._perm = ids.data.astype(np.int64) - 1          # csc slot -> COO entry
        self._indices, self._indptr, self._n = ids.indices.copy(), ids.indptr.copy(), pattern.n
        self.outputs = getattr(assemble, "outputs", None)

    def __call__(self, U, xi_prev):                      # plain assembly (sensitivities etc.)
        return self._assemble(U, xi_prev)

    def enforced(self, U, xi_prev, t):
        import torch
        R, K_data, xi = self._assemble.device(U, xi_prev)
        Ud = torch.from_numpy(np.ascontiguousarray(U)).to(self._dev)
        vals = torch.from_numpy(np.ascontiguousarray(self._bcs.values(t), dtype=np.float64)).to(self._dev)
        r, K_emb = self._plan.apply(K_data, R, Ud, vals)
        data = K_emb.cpu().numpy()[self._perm]
        return r.cpu().numpy(), sp.csc_matrix((data, self._indices, self._indptr), shape=(self._n, self._n)), xi
