"""CPU restatement of the FE element-block assembly (COUPLED mode).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs, never by the
product package.

Follows, vectorised over elements with NumPy:
  cmad/global_residuals/interpolation.py:49-55      grad_u = U_e^T grad_N
  cmad/global_residuals/global_residual.py:361-395  per-IP evaluator: local Newton from
                                                    xi_prev, R, IFT-corrected dR/dU
  cmad/global_residuals/small_disp_equilibrium.py:112-118  R = (grad_N @ sigma) * w * dv
  cmad/fem/assembly.py:416-535   scan over IPs: R_e, K_e summed in IP order, xi stacked
  cmad/fem/assembly.py:661-732   gather U, scatter-add R, emit COO ``vals``
  cmad/fem/assembly.py:906-909   COO dedup (segment sum through coo_dedup_scatter)
The per-point solve, stress and consistent tangent come from the dual-number
C++ oracle (``oracle_c``), i.e. are AD-derived like the reference's.  The
line-by-line torch-AD restatement of one element (``cmad_oracle.coupled_element``)
cross-checks this file in tests/.
"""
from __future__ import annotations

import numpy as np

from . import oracle_c

# symmetric component (xx,xy,xz,yy,yz,zz) of the 3x3 entry (i, j)
_V = np.array([[0, 1, 2], [1, 3, 4], [2, 4, 5]])


def full_tangent(dsig_deps: np.ndarray) -> np.ndarray:
    """``Dfull[n, j, i, k, l] = d sigma_ji / d grad_u_kl`` from the symmetric-strain
    tangent ``dsig_deps (36, n)`` (a*6+b): a grad_u entry (k,l), k != l, moves the
    symmetric component by 1/2."""
    n = dsig_deps.shape[1]
    D = dsig_deps.reshape(6, 6, n)
    half = np.array([1.0, 0.5, 0.5, 1.0, 0.5, 1.0])
    Dh = D * half[None, :, None]
    return np.moveaxis(Dh[_V][:, :, _V], -1, 0)          # (n, 3, 3, 3, 3) [j, i, k, l]


def assemble_block(prob: oracle_c.OracleProblem, elem_eq, U, xi_prev, grad_N, det, quad_w,
                   want_K: bool = True, nthreads: int = 0) -> dict:
    """One COUPLED element block.  Inputs in the reference's layouts: ``elem_eq
    (n_e, n_b*3)``, ``U (n_dofs,)``, ``xi_prev (n_e, n_ip, 7)``, ``grad_N (n_e, n_ip,
    n_b, 3)``, ``det (n_e, n_ip)``, ``quad_w (n_ip,)``.  ``prob`` must be described
    with ``strain_comps=9``.  Returns ``R_elem (n_e, n_b*3)``, ``K_elem (n_e, n_b*3,
    n_b*3)``, ``xi``, ``sigma (n_e, n_ip, 6)``, ``iters``, ``flags``, ``R (n_dofs,)``."""
    elem_eq = np.asarray(elem_eq, dtype=np.int64)
    n_e, n_ip, n_b, _ = grad_N.shape
    assert int(prob.cfg[7]) == 9
    U_e = np.asarray(U)[elem_eq].reshape(n_e, n_b, 3)                    # assembly.py:121-139
    R_e = np.zeros((n_e, n_b, 3))
    K_e = np.zeros((n_e, n_b, 3, n_b, 3)) if want_K else None
    xi = np.zeros((n_e, n_ip, 7)); sigma = np.zeros((n_e, n_ip, 6))
    iters = np.zeros((n_e, n_ip), dtype=np.int32); flags = np.zeros((n_e, n_ip), dtype=np.int32)
    want = ("xi", "sigma", "iters", "flags") + (("dsig_deps",) if want_K else ())
    for ip in range(n_ip):                                               # lax.scan over IPs
        gN = grad_N[:, ip]                                               # (n_e, n_b, 3)
        gu = np.einsum("eak,eaj->ekj", U_e, gN)                          # grad_u[k, j]
        r = oracle_c.mp_update(prob, xi_prev[:, ip].T.copy(), gu.reshape(n_e, 9).T.copy(),
                               want=want, nthreads=nthreads)
        xi[:, ip] = r["xi"].T; sigma[:, ip] = r["sigma"].T
        iters[:, ip] = r["iters"]; flags[:, ip] = r["flags"]
        sig33 = r["sigma"][_V]                                           # (3, 3, n_e)
        wdv = quad_w[ip] * det[:, ip]
        R_e += np.einsum("eaj,jie->eai", gN, sig33) * wdv[:, None, None]
        if want_K:
            Df = full_tangent(r["dsig_deps"])
            K_e += np.einsum("eaj,ejikl,ebl->eaibk", gN, Df, gN) * wdv[:, None, None, None, None]
    out = {"R_elem": R_e.reshape(n_e, n_b * 3), "xi": xi, "sigma": sigma, "iters": iters, "flags": flags}
    if want_K:
        out["K_elem"] = K_e.reshape(n_e, n_b * 3, n_b * 3)
    R = np.zeros(np.asarray(U).shape[0])
    np.add.at(R, elem_eq.reshape(-1), out["R_elem"].reshape(-1))         # assembly.py:715-720
    out["R"] = R
    return out


def assemble_block_mixed(prob: oracle_c.OracleProblem, elem_eq, elem_eq_p, U, xi_prev, grad_N, N, det,
                         quad_w, h, stab_mult: float = 1.0, want_K: bool = True, nthreads: int = 0) -> dict:
    """One COUPLED element block of the MIXED u-p formulation
    (cmad/global_residuals/small_disp_equilibrium.py:87-111), per IP:
      sigma = dev(cauchy(xi)) - p I,                      R_u = (grad_N @ sigma) w dv
      R_p = (-(p + hydro)/kappa N - tau grad_N @ grad p) w dv,  hydro = kappa tr(eps),
      tau = mult * 0.5 h^2 / mu
    and the four tangent blocks dR_r/dU_s (IFT-corrected through the local Newton, which
    only sees grad_u).  ``U`` holds the block-major (u, p) dofs; ``N (n_ip, n_b)``; ``h (n_e,)``.
    Returns R_u (n_e, 3n_b), R_p (n_e, n_b), K_uu, K_up, K_pu, K_pp, xi, R (n_dofs,)."""
    elem_eq = np.asarray(elem_eq, dtype=np.int64); elem_eq_p = np.asarray(elem_eq_p, dtype=np.int64)
    n_e, n_ip, n_b, _ = grad_N.shape
    assert int(prob.cfg[7]) == 9
    U = np.asarray(U)
    U_e = U[elem_eq].reshape(n_e, n_b, 3); p_e = U[elem_eq_p]
    lam, mu = oracle_c.lame(int(prob.cfg[2]), float(prob.mat[0]), float(prob.mat[1]))[:2]
    kappa = lam + 2.0 * mu / 3.0                                         # elastic_constants.py:41-43
    tau = stab_mult * 0.5 * np.asarray(h) ** 2 / mu
    R_u = np.zeros((n_e, n_b, 3)); R_p = np.zeros((n_e, n_b))
    K_uu = np.zeros((n_e, n_b, 3, n_b, 3)); K_up = np.zeros((n_e, n_b, 3, n_b))
    K_pu = np.zeros((n_e, n_b, n_b, 3)); K_pp = np.zeros((n_e, n_b, n_b))
    xi = np.zeros((n_e, n_ip, 7)); sigma_ip = np.zeros((n_e, n_ip, 6))
    I3 = np.eye(3)
    want = ("xi", "sigma") + (("dsig_deps",) if want_K else ())
    for ip in range(n_ip):
        gN, Nip = grad_N[:, ip], N[ip]
        gu = np.einsum("eak,eaj->ekj", U_e, gN)
        r = oracle_c.mp_update(prob, xi_prev[:, ip].T.copy(), gu.reshape(n_e, 9).T.copy(), want=want,
                               nthreads=nthreads)
        xi[:, ip] = r["xi"].T; sigma_ip[:, ip] = r["sigma"].T
        sig = np.moveaxis(r["sigma"][_V], -1, 0)                         # (n_e, 3, 3) cauchy
        dev = sig - (np.trace(sig, axis1=1, axis2=2) / 3.0)[:, None, None] * I3      # dev_cauchy :323-329
        p = p_e @ Nip
        sigma = dev - p[:, None, None] * I3
        wdv = quad_w[ip] * det[:, ip]
        R_u += np.einsum("eaj,eji->eai", gN, sigma) * wdv[:, None, None]
        hydro = kappa * np.trace(gu, axis1=1, axis2=2)                   # hydro_cauchy :332-339
        grad_p = np.einsum("ea,eaj->ej", p_e, gN)
        R_p += (-(p + hydro)[:, None] / kappa * Nip[None, :]
                - tau[:, None] * np.einsum("eaj,ej->ea", gN, grad_p)) * wdv[:, None]
        if want_K:
            Df = full_tangent(r["dsig_deps"])                            # [e, j, i, k, l]
            tr = np.einsum("emmkl->ekl", Df) / 3.0
            Ddev = Df - I3[None, :, :, None, None] * tr[:, None, None, :, :]
            K_uu += np.einsum("eaj,ejikl,ebl->eaibk", gN, Ddev, gN) * wdv[:, None, None, None, None]
            K_up += -np.einsum("eai,b->eaib", gN, Nip) * wdv[:, None, None, None]
            K_pu += -np.einsum("a,ebk->eabk", Nip, gN) * wdv[:, None, None, None]
            K_pp += (-np.einsum("a,b->ab", Nip, Nip)[None] / kappa
                     - tau[:, None, None] * np.einsum("eaj,ebj->eab", gN, gN)) * wdv[:, None, None]
    out = {"R_u": R_u.reshape(n_e, n_b * 3), "R_p": R_p, "xi": xi, "sigma": sigma_ip}
    if want_K:
        out.update(K_uu=K_uu.reshape(n_e, 3 * n_b, 3 * n_b), K_up=K_up.reshape(n_e, 3 * n_b, n_b),
                   K_pu=K_pu.reshape(n_e, n_b, 3 * n_b), K_pp=K_pp)
    R = np.zeros(U.shape[0])
    np.add.at(R, elem_eq.reshape(-1), out["R_u"].reshape(-1))
    np.add.at(R, elem_eq_p.reshape(-1), R_p.reshape(-1))
    out["R"] = R
    return out


def cauchy_at_ips(prob_eval: oracle_c.OracleProblem, elem_eq, U, xi, grad_N) -> np.ndarray:
    """``evaluate_cauchy_at_ips`` for a COUPLED block (cmad/fem/postprocess.py:35-185):
    ``model.cauchy(xi, ., params, U_ip, .)`` at the stored state; ``prob_eval`` described with
    ``strain_comps=9, max_iters=0``.  Returns ``(n_e, n_ip, 6)``."""
    elem_eq = np.asarray(elem_eq, dtype=np.int64)
    n_e, n_ip, n_b, _ = grad_N.shape
    U_e = np.asarray(U)[elem_eq].reshape(n_e, n_b, 3)
    out = np.zeros((n_e, n_ip, 6))
    for ip in range(n_ip):
        gu = np.einsum("eak,eaj->ekj", U_e, grad_N[:, ip]).reshape(n_e, 9).T.copy()
        x = xi[:, ip].T.copy()
        out[:, ip] = oracle_c.mp_update(prob_eval, x, gu, xi_init=x, want=("sigma",))["sigma"].T
    return out


def embedded_system(rows, cols, n, K_data, R, U, presc_idx, presc_vals):
    """``_embedded_bc_enforce`` + ``_embedded_residual`` (cmad/fem/sparse_solve.py:1058-1174) on
    a deduplicated COO pattern, restated with plain NumPy loops over the entries:
    returns ``(r, K_emb_data)`` on the same pattern (prescribed diagonal = assembled K_ii)."""
    rows = np.asarray(rows); cols = np.asarray(cols); K = np.asarray(K_data, dtype=np.float64)
    p_mask = np.zeros(n, dtype=bool); p_mask[presc_idx] = True
    keep = ~(p_mask[rows] | p_mask[cols])                          # :1146-1147
    K_ii_full = np.zeros(n); np.add.at(K_ii_full, rows, K * (rows == cols))      # :1149-1152
    K_ii = K_ii_full[presc_idx]
    K_emb = K * keep
    diag_presc = (rows == cols) & p_mask[rows]
    K_emb[diag_presc] = K[diag_presc]                              # appended (presc, presc) entries, deduplicated
    inc = np.zeros(n); inc[presc_idx] = np.asarray(presc_vals) - np.asarray(U)[presc_idx]      # :1173-1175
    r = np.asarray(R, dtype=np.float64).copy()
    np.add.at(r, rows, K * inc[cols])
    r[presc_idx] = K_ii * (np.asarray(U)[presc_idx] - np.asarray(presc_vals))                 # :1177-1179
    return r, K_emb


def coo_dedup_sum(vals: np.ndarray, scatter: np.ndarray, n_unique: int) -> np.ndarray:
    """``zeros(n_unique).at[coo_dedup_scatter].add(vals)`` (assembly.py:906-909)."""
    out = np.zeros(n_unique)
    np.add.at(out, scatter, vals)
    return out


def block_jvp(prob_eval: oracle_c.OracleProblem, elem_eq, U, xi_prev, xi_state, grad_N, det, quad_w,
              dp, dxi_prev=None, nthreads: int = 0, dU=None, _want_dsigma: bool = False) -> dict:
    """Forward sensitivities of a COUPLED block at the converged state, at fixed ``U``
    (what jax.jvp pushes through the custom_jvp rule of make_newton_solve,
    cmad/models/nonlinear_solver.py:158-171, inside cmad/fem/nonlinear_solver.py:490-537):
    ``dxi = -A^{-1}(dC/dp dp + dC/dxi_prev dxi_prev)``, ``dsigma = dcauchy/dxi dxi +
    dcauchy/dp dp``, ``dR_e = sum_ip gradN dsigma w dv``.  ``prob_eval`` must be described
    with ``strain_comps=9, max_iters=0`` (evaluation at ``xi_init = xi_state``) and the
    active parameters; every derivative block comes from the dual-number C++ oracle."""
    elem_eq = np.asarray(elem_eq, dtype=np.int64)
    n_e, n_ip, n_b, _ = grad_N.shape
    na = len(prob_eval.active_pid)
    dp = np.asarray(dp, dtype=np.float64).reshape(na)
    U_e = np.asarray(U)[elem_eq].reshape(n_e, n_b, 3)
    dU_e = None if dU is None else np.asarray(dU)[elem_eq].reshape(n_e, n_b, 3)
    dR = np.zeros((n_e, n_b, 3)); dxi = np.zeros((n_e, n_ip, 7)); dsig = []
    for ip in range(n_ip):
        gN = grad_N[:, ip]
        gu = np.einsum("eak,eaj->ekj", U_e, gN).reshape(n_e, 9).T.copy()
        want = ("dC_dxi", "dC_dxi_prev", "dC_dp", "dsig_dxi", "dsig_dp") + \
            (("dxi_deps", "dsig_deps") if dU is not None else ())
        r = oracle_c.mp_update(prob_eval, xi_prev[:, ip].T.copy(), gu, xi_init=xi_state[:, ip].T.copy(),
                               want=want, nthreads=nthreads)
        A = np.moveaxis(r["dC_dxi"].reshape(7, 7, n_e), 2, 0)
        B = np.moveaxis(r["dC_dxi_prev"].reshape(7, 7, n_e), 2, 0)
        rhs = np.zeros((n_e, 7))
        if na:
            rhs += np.einsum("rce,c->er", r["dC_dp"].reshape(7, na, n_e), dp)
        if dxi_prev is not None:
            rhs += np.einsum("erc,ec->er", B, dxi_prev[:, ip])
        dx = -np.linalg.solve(A, rhs[:, :, None])[:, :, 0]
        ds = np.einsum("ace,ec->ae", r["dsig_dxi"].reshape(6, 7, n_e), dx)
        if na:
            ds += np.einsum("ace,c->ae", r["dsig_dp"].reshape(6, na, n_e), dp)
        if dU is not None:
            # displacement direction: the IFT blocks dxi/deps and the TOTAL dsigma/deps of
            # the oracle (symmetric strain components, both tensor entries moving)
            dg = np.einsum("eak,eaj->ekj", dU_e, gN)
            de = 0.5 * (dg + np.swapaxes(dg, 1, 2))
            de6 = np.stack([de[:, 0, 0], de[:, 0, 1], de[:, 0, 2], de[:, 1, 1], de[:, 1, 2], de[:, 2, 2]])
            dx = dx + np.einsum("rbe,be->er", r["dxi_deps"].reshape(7, 6, n_e), de6)
            ds = ds + np.einsum("abe,be->ae", r["dsig_deps"].reshape(6, 6, n_e), de6)
        dxi[:, ip] = dx
        dsig.append(ds)
        wdv = quad_w[ip] * det[:, ip]
        dR += np.einsum("eaj,jie->eai", gN, ds[_V]) * wdv[:, None, None]
    out = {"R_elem": dR.reshape(n_e, n_b * 3), "xi": dxi}
    if _want_dsigma:
        out["dsigma"] = dsig                     # per IP: (6, n_e) direction of the global cauchy
    return out


def block_jvp_mixed(prob_eval: oracle_c.OracleProblem, elem_eq, elem_eq_p, U, xi_prev, xi_state, grad_N, N, det,
                    quad_w, h, dp, dxi_prev=None, stab_mult: float = 1.0, nthreads: int = 0,
                    dU=None) -> dict:
    """Forward sensitivities of a MIXED u-p COUPLED block at the converged state (the mixed
    counterpart of :func:`block_jvp`): what jax.jvp pushes through the two residual blocks of
    cmad/global_residuals/small_disp_equilibrium.py:87-111 inside cmad/fem/nonlinear_solver.py:
    490-537.  The local Newton only sees grad_u, so ``dxi`` and ``d cauchy`` are those of the
    displacement form; then
      dR_u = sum_ip gradN (dev(d cauchy) - dp_ip I) w dv,
      dR_p = sum_ip ( -(dp_ip + d hydro)/kappa N + (p + hydro) dkappa/kappa^2 N
                      - d tau gradN.grad p - tau gradN.grad dp ) w dv,
    hydro = kappa tr(eps) (so the dkappa terms of hydro cancel except p dkappa/kappa^2),
    tau = mult h^2/(2 mu), d tau = -tau dmu/mu, with (dlam, dmu) the direction of the Lame
    constants for ``dp`` (dual-number Jacobian of elastic_constants.py:54-104)."""
    elem_eq = np.asarray(elem_eq, dtype=np.int64); elem_eq_p = np.asarray(elem_eq_p, dtype=np.int64)
    n_e, n_ip, n_b, _ = grad_N.shape
    U = np.asarray(U)
    base = block_jvp(prob_eval, elem_eq, U, xi_prev, xi_state, grad_N, det, quad_w, dp, dxi_prev,
                     nthreads=nthreads, dU=dU, _want_dsigma=True)
    lj = oracle_c.lame(int(prob_eval.cfg[2]), float(prob_eval.mat[0]), float(prob_eval.mat[1]))
    lam, mu = lj[:2]
    kappa = lam + 2.0 * mu / 3.0
    dlam = dmu = 0.0
    for c, pid in enumerate(np.asarray(prob_eval.active_pid)):
        if pid in (oracle_c.PID["EL0"], oracle_c.PID["EL1"]):
            k = int(pid) - oracle_c.PID["EL0"]
            dlam += lj[2 + k] * float(np.asarray(dp)[c]); dmu += lj[4 + k] * float(np.asarray(dp)[c])
    dkappa = dlam + 2.0 * dmu / 3.0
    tau = stab_mult * 0.5 * np.asarray(h) ** 2 / mu
    dtau = -tau * dmu / mu
    U_e = U[elem_eq].reshape(n_e, n_b, 3); p_e = U[elem_eq_p]
    dU_e = np.zeros_like(U_e) if dU is None else np.asarray(dU)[elem_eq].reshape(n_e, n_b, 3)
    dp_e = np.zeros_like(p_e) if dU is None else np.asarray(dU)[elem_eq_p]
    I3 = np.eye(3)
    dR_u = np.zeros((n_e, n_b, 3)); dR_p = np.zeros((n_e, n_b))
    for ip in range(n_ip):
        gN, Nip = grad_N[:, ip], N[ip]
        ds = np.moveaxis(base["dsigma"][ip][_V], -1, 0)                   # (n_e, 3, 3)
        dpi = dp_e @ Nip
        dsm = ds - (np.trace(ds, axis1=1, axis2=2) / 3.0)[:, None, None] * I3 - dpi[:, None, None] * I3
        wdv = quad_w[ip] * det[:, ip]
        dR_u += np.einsum("eaj,eji->eai", gN, dsm) * wdv[:, None, None]
        p = p_e @ Nip
        tre = np.trace(np.einsum("eak,eaj->ekj", U_e, gN), axis1=1, axis2=2)
        dtre = np.trace(np.einsum("eak,eaj->ekj", dU_e, gN), axis1=1, axis2=2)
        hydro, dhydro = kappa * tre, dkappa * tre + kappa * dtre
        grad_p = np.einsum("ea,eaj->ej", p_e, gN); grad_dp = np.einsum("ea,eaj->ej", dp_e, gN)
        dR_p += ((-(dpi + dhydro) / kappa + (p + hydro) * dkappa / kappa ** 2)[:, None] * Nip[None, :]
                 - dtau[:, None] * np.einsum("eaj,ej->ea", gN, grad_p)
                 - tau[:, None] * np.einsum("eaj,ej->ea", gN, grad_dp)) * wdv[:, None]
    return {"R_elem": dR_u.reshape(n_e, n_b * 3), "R_p_elem": dR_p, "xi": base["xi"]}
