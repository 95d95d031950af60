"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN, UNMODIFIED SOURCE
(`/root/reference/cmad/...`) for the hot path.

The reference is pure Python on JAX; JAX is not installable in the build
container.  `tests/golden/jaxshim/` provides the slice of the JAX API the path
uses (arrays, forward-mode AD transforms, eager control flow, pytrees) on NumPy
fp64, so the reference's constitutive code - `SmallElasticPlastic._residual_fn`,
the effective stresses, hardening, `cond_residual`, `make_newton_solve` and its
`custom_jvp` IFT rule, `line_search`, the imperative `newton_solve`, `Model`'s
AD products, `Parameters`, `MPAdjointObjective` / `MPDirectObjective`,
`Calibration`, and the FE per-IP evaluator of `GlobalResidual._for_model_coupled`
- runs line by line as written.  Only the arithmetic library differs from a real
JAX run (NumPy/LAPACK instead of XLA), i.e. rounding-level.

Run in the build container only (needs /root/reference):

    python tests/golden/make_reference_golden.py [--jobs 8]

Writes tests/golden/ref_*.npz; the fixtures are committed, `/root/reference` is
never read at test time.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def _enter_reference():
    sys.path[:0] = [os.path.join(HERE, "jaxshim"), REFERENCE, ROOT]
    for m in ("netCDF4", "gmsh", "pyamg", "matplotlib", "matplotlib.pyplot", "sympy", "jsonschema"):
        try:
            __import__(m)
        except ImportError:
            sys.modules[m] = types.ModuleType(m)


_enter_reference()

import numpy as np  # noqa: E402

import jax  # noqa: E402  (the shim)
from jax import _core  # noqa: E402
from cmad.models.deriv_types import DerivType  # noqa: E402
from cmad.models.effective_stress import conventional_effective_stress_fun  # noqa: E402
from cmad.models.elastic_stress import isotropic_linear_elastic_stress  # noqa: E402
from cmad.models.global_fields import GlobalFieldsAtPoint, mp_U_from_F  # noqa: E402
from cmad.models.hardening import combined_hardening_fun, get_hardening_funs  # noqa: E402
from cmad.models.nonlinear_solver import make_newton_solve, newton_solve  # noqa: E402
from cmad.models.small_elastic_plastic import SmallElasticPlastic, compute_yield_fun_and_normal  # noqa: E402
from cmad.objectives.mp_objective import MPAdjointObjective, MPDirectAdjointObjective, MPDirectObjective  # noqa: E402
from cmad.parameters.parameters import Parameters  # noqa: E402
from cmad.qois.calibration import Calibration  # noqa: E402
from functools import partial  # noqa: E402

from cmad_b200 import synthetic  # noqa: E402  (input generator only: the inputs are stored in the fixture)

UP = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]


sys.path.insert(0, HERE)
from materials import NEWTON, active_all_scalars, const_like, material, objective_trees  # noqa: E402


def parameters(values):
    return Parameters(values, active_all_scalars(values), const_like(values, None))


def vec6(T):
    T = np.asarray(T)
    return np.array([T[i, j] for i, j in UP])


def yield_fun_of(model, values):
    kind = next(iter(values["plastic"]["effective stress"]))
    return partial(compute_yield_fun_and_normal, def_type=0,
                   elastic_stress=isotropic_linear_elastic_stress,
                   effective_stress=conventional_effective_stress_fun(kind),
                   hardening=partial(combined_hardening_fun, hardening_funs=get_hardening_funs()),
                   uniaxial_stress_idx=0, is_complex=False)


# --------------------------------------------------------------------------- #
#  A. traced Newton (make_newton_solve + IFT rule), batched over points       #
# --------------------------------------------------------------------------- #
def _traced_chunk(job):
    kind, newton_key, lo, hi, steps, seed, scale, with_tangent = job
    values = material(kind)
    P = parameters(values)
    model = SmallElasticPlastic(P)
    solve = make_newton_solve(model._residual, **NEWTON[newton_key])
    yf = yield_fun_of(model, values)
    d, d2, a = synthetic.path_params(seed, lo, hi - lo, diag_only=kind.startswith("hosford"))
    n = hi - lo
    nP = P.num_params
    out = {k: [] for k in ("grad_u", "xi_prev", "xi", "iters", "flags", "sigma", "dsig_dgradu",
                           "dxi_dgradu", "dC_dp", "dC_dxi", "dC_dxi_prev", "dxi_dp", "dsig_dp")}
    rng = np.random.default_rng(1000 + lo)
    xi_prev = [[np.zeros(6), np.zeros(1)] for _ in range(n)]
    for t in steps:
        e6 = synthetic.strain_at_step(d, d2, a * scale, t)                    # (6, n)
        rec = {k: [] for k in out}
        for i in range(n):
            e = e6[:, i]
            w = rng.normal(size=3) * 1e-3                                     # rigid-rotation part: must not matter
            gu = np.array([[e[0], e[1] + w[0], e[2] + w[1]],
                           [e[1] - w[0], e[3], e[4] + w[2]],
                           [e[2] - w[1], e[4] - w[2], e[5]]])
            U = mp_U_from_F(np.eye(3) + gu)
            xp = xi_prev[i]
            params = P.values
            xi = solve(xp, params, U, U)
            iters = int(_core.WHILE_LOG[-1][1][0])
            xi_np = [np.asarray(x) for x in xi]
            _, f0, _ = yf(xp, xp, params, U, U)
            _, f1, _ = yf(xi_np, xp, params, U, U)
            pl = lambda f: bool(f > 1e-14 or abs(f) < 1e-14)                  # noqa: E731  paths.py:26
            flags = (1 if pl(float(f0)) else 0) | (2 if pl(float(f1)) else 0)
            sig = np.asarray(model.cauchy(xi_np, xp, params, U, U))
            rec["grad_u"].append(gu.reshape(9)); rec["xi_prev"].append(np.concatenate(xp))
            rec["xi"].append(np.concatenate(xi_np)); rec["iters"].append(iters); rec["flags"].append(flags)
            rec["sigma"].append(vec6(sig))
            # Model AD products at the converged state (model.py:126-133, parameters.py:368-377)
            jac = model._jacobian
            rec["dC_dxi"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI](xi_np, xp, params, U, U)]))
            rec["dC_dxi_prev"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI_PREV](xi_np, xp, params, U, U)]))
            dcdp = jac[DerivType.DPARAMS](xi_np, xp, params, U, U)
            flat = [np.asarray(x).reshape(7, -1) for x in jax.tree_util.tree_leaves(dcdp)]
            rec["dC_dp"].append(np.hstack(flat))                              # (7, P) all leaves, flatten order
            if with_tangent:
                def state_and_stress(U_, p_):
                    x = solve(xp, p_, U_, U)
                    return jax.numpy.concatenate([jax.numpy.ravel(b) for b in x]), model.cauchy(x, xp, p_, U_, U)
                (dx_dU, ds_dU) = jax.jacfwd(state_and_stress, argnums=0)(U, params)
                rec["dxi_dgradu"].append(np.asarray(dx_dU.grad_fields["u"]).reshape(7, 9))
                rec["dsig_dgradu"].append(np.asarray(ds_dU.grad_fields["u"]).reshape(9, 9))
                (dx_dp, ds_dp) = jax.jacfwd(state_and_stress, argnums=1)(U, params)
                rec["dxi_dp"].append(np.hstack([np.asarray(x).reshape(7, -1) for x in jax.tree_util.tree_leaves(dx_dp)]))
                rec["dsig_dp"].append(np.hstack([np.asarray(x).reshape(9, -1) for x in jax.tree_util.tree_leaves(ds_dp)]))
            xi_prev[i] = xi_np
        for k in out:
            if rec[k]:
                out[k].append(np.array(rec[k]))
    return {k: np.array(v) for k, v in out.items() if v}, nP       # each (steps, n, ...)


def _partials_job(job):
    """The raw AD products of `Model.__init__` (cmad/models/model.py:121-160) that the traced fixture
    does not hold: dC/dU, dC/dU_prev (jacfwd over the GlobalFieldsAtPoint argument) and the partial
    derivatives of `cauchy` with respect to xi, xi_prev and the parameters, at converged states."""
    kind, newton_key, n, steps, seed = job
    values = material(kind)
    P = parameters(values)
    model = SmallElasticPlastic(P)
    solve = make_newton_solve(model._residual, **NEWTON[newton_key])
    d, d2, a = synthetic.path_params(seed, 0, n, diag_only=kind.startswith("hosford"))
    out = {k: [] for k in ("grad_u", "xi_prev", "xi", "flags", "dC_dU", "dC_dU_prev", "dsig_dxi", "dsig_dxi_prev",
                           "dsig_dp", "dsig_dU")}
    yf = yield_fun_of(model, values)
    xi_prev = [[np.zeros(6), np.zeros(1)] for _ in range(n)]
    for t in steps:
        e6 = synthetic.strain_at_step(d, d2, a, t)
        for i in range(n):
            e = e6[:, i]
            gu = np.array([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]])
            U = mp_U_from_F(np.eye(3) + gu)
            xp, params = xi_prev[i], P.values
            xi = [np.asarray(x) for x in solve(xp, params, U, U)]
            _, f1, _ = yf(xi, xp, params, U, U)
            out["flags"].append(2 if (float(f1) > 1e-14 or abs(float(f1)) < 1e-14) else 0)
            jac = model._jacobian
            out["grad_u"].append(gu.reshape(9)); out["xi_prev"].append(np.concatenate(xp)); out["xi"].append(np.concatenate(xi))
            out["dC_dU"].append(np.asarray(jac[DerivType.DU](xi, xp, params, U, U).grad_fields["u"]).reshape(7, 9))
            out["dC_dU_prev"].append(np.asarray(jac[DerivType.DU_PREV](xi, xp, params, U, U).grad_fields["u"]).reshape(7, 9))
            out["dsig_dxi"].append(np.hstack([np.asarray(b).reshape(9, -1) for b in model.dcauchy[0](xi, xp, params, U, U)]))
            out["dsig_dxi_prev"].append(np.hstack([np.asarray(b).reshape(9, -1) for b in model.dcauchy[1](xi, xp, params, U, U)]))
            out["dsig_dp"].append(np.hstack([np.asarray(x).reshape(9, -1)
                                             for x in jax.tree_util.tree_leaves(model.dcauchy[2](xi, xp, params, U, U))]))
            dsdU = jax.jacfwd(model.cauchy, argnums=DerivType.DU)(xi, xp, params, U, U)
            out["dsig_dU"].append(np.asarray(dsdU.grad_fields["u"]).reshape(9, 9))
            xi_prev[i] = xi
    res = {k: np.array(v) for k, v in out.items()}
    res["param_names"] = np.array(P._names)
    res["param_sizes"] = np.array(P.flat_param_sizes)
    return res


def traced(pool, kind, newton_key, n, steps, seed=22, scale=1.0, chunk=4, with_tangent=True):
    jobs = [(kind, newton_key, lo, min(lo + chunk, n), steps, seed, scale, with_tangent)
            for lo in range(0, n, chunk)]
    parts = pool.map(_traced_chunk, jobs)
    res = {k: np.concatenate([p[0][k] for p in parts], axis=1) for k in parts[0][0]}
    values = material(kind)
    P = parameters(values)
    res["param_names"] = np.array(P._names)
    res["param_sizes"] = np.array(P.flat_param_sizes)
    res["steps"] = np.array(steps)
    return res


# --------------------------------------------------------------------------- #
#  B. imperative Newton through the Model object (newton_solve -> solver.json) #
# --------------------------------------------------------------------------- #
def two_leg_F(seed, nsteps, scale=1.0, diag_only=False):
    d, d2, a = synthetic.path_params(seed, 0, 1, diag_only=diag_only)
    F = np.repeat(np.eye(3)[:, :, None], nsteps + 1, axis=2)
    for t in range(1, nsteps + 1):
        e = synthetic.strain_at_step(d, d2, a * scale, int(round(t * 100 / nsteps)))[:, 0]
        F[:, :, t] += np.array([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]])
    return F


def _imperative_job(job):
    kind, F = job
    values = material(kind)
    model = SmallElasticPlastic(parameters(values))
    N = F.shape[2] - 1
    xi = np.zeros((N + 1, 7)); sig = np.zeros((N + 1, 6)); it = np.zeros(N + 1, int); cn = np.zeros(N + 1)
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        model.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
        it[step], cn[step] = newton_solve(model)                  # mp_objective.py:83, defaults 10/1e-14/1e-14
        xi[step] = np.concatenate([np.asarray(b) for b in model.xi()])
        model.seed_none()
        model.evaluate_cauchy()
        sig[step] = vec6(model.Sigma())
        model.advance_xi()
    return dict(F=F, xi=xi, sigma=sig, iters=it, cnorm=cn)


# --------------------------------------------------------------------------- #
#  C. MP adjoint / direct objectives with the Calibration QoI                 #
# --------------------------------------------------------------------------- #
def _objective_job(job):
    kind, scaled, F, weight = job[:4]
    values, act, tr = objective_trees(kind, scaled)
    P = Parameters(values, act, tr)
    if len(job) > 4 and job[4] == "rate":       # the rate form under the same objectives
        from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic
        model = SmallRateElasticPlastic(P)
    else:
        model = SmallElasticPlastic(P)
    N = F.shape[2] - 1
    # data: the stress history at the "true" parameters (test_J2_fd_checks.py:21-47)
    data = np.zeros((3, 3, N + 1))
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        model.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
        newton_solve(model)
        model.seed_none(); model.evaluate_cauchy()
        data[:, :, step] = model.Sigma().copy()
        model.advance_xi()
    qoi = Calibration(model, data, weight)
    true_vals = P.flat_active_values(False)
    offset = 1.1 * true_vals                                      # test_J2_fd_checks.py:326
    P.set_active_values_from_flat(offset, False)
    x = P.flat_active_values(True)                                # canonical coordinates of the offset point
    res = {}
    for name, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        P.set_active_values_from_flat(offset, False)
        if len(job) > 4 and job[4] == "rate":
            # MPObjective.__init__ (mp_objective.py:46) stores the model's CURRENT xi as the step-0
            # state of the adjoint pass.  After the data run above that is the final state of the
            # data history, not the initial one the forward pass starts from; harmless for
            # SmallElasticPlastic with an elastic first step (nothing there depends on xi_prev), but
            # the rate form's dC/dp reads sig_prev: build the objective on a model at its initial state.
            model.set_xi_to_init_vals()
        J, g = ctor(qoi, F).evaluate(x)
        res[f"J_{name}"], res[f"grad_{name}"] = float(J), np.asarray(g, float)
    res.update(F=F, data=data, weight=weight, x_canonical=x, active_native=offset,
               active_idx=np.asarray(P.active_idx), param_names=np.array(P._names))
    return res



def _hessian_job(job):
    """MPDirectAdjointObjective (mp_objective.py:218-343): J, gradient and Hessian in
    canonical coordinates through Model.evaluate_hessians / qoi.evaluate_hessians."""
    kind, scaled, F, weight = job[:4]
    dt_name = job[4] if len(job) > 4 else "FULL_3D"
    from cmad.models.deformation_types import DefType
    Model_ = SmallElasticPlastic
    if len(job) > 5 and job[5] == "rate":       # the rate form (small_rate_elastic_plastic.py)
        from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic as Model_
    values, act, tr = objective_trees(kind, scaled)
    P = Parameters(values, act, tr)
    model = Model_(P, def_type=DefType[dt_name])
    N = F.shape[2] - 1
    data = np.zeros((3, 3, N + 1))
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        model.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
        newton_solve(model)
        model.seed_none(); model.evaluate_cauchy()
        data[:, :, step] = model.Sigma().copy()
        model.advance_xi()
    true_vals = P.flat_active_values(False)
    offset = 1.1 * true_vals
    # a FRESH parameters / model pair for the objective, as a run that reads its data from file
    # has (the data-generation model above is not reused)
    P = Parameters(*objective_trees(kind, scaled))
    model = Model_(P, def_type=DefType[dt_name])
    qoi = Calibration(model, data, weight)
    P.set_active_values_from_flat(offset, False)
    x = P.flat_active_values(True)
    r = MPDirectAdjointObjective(qoi, F).evaluate(x)
    return dict(J=float(r.J), grad=np.asarray(r.grad, float), hessian=np.asarray(r.hessian, float),
                F=F, data=data, weight=weight, x_canonical=x, active_native=offset,
                active_idx=np.asarray(P.active_idx), param_names=np.array(P._names))


def _hessian_jvp_job(job):
    """MPJVPObjective (mp_jvp_objective.py:14-80): jax.value_and_grad / jax.hessian of the whole
    time loop through make_newton_solve and its custom_jvp rule - the reference's OTHER Hessian
    strategy, with no hand-assembled blocks."""
    from cmad.objectives.mp_jvp_objective import MPJVPObjective
    kind, scaled, F, weight = job
    values, act, tr = objective_trees(kind, scaled)
    P = Parameters(values, act, tr)
    model = SmallElasticPlastic(P)
    N = F.shape[2] - 1
    data = np.zeros((3, 3, N + 1))
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        model.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
        newton_solve(model)
        model.seed_none(); model.evaluate_cauchy()
        data[:, :, step] = model.Sigma().copy()
        model.advance_xi()
    offset = 1.1 * P.flat_active_values(False)
    P = Parameters(*objective_trees(kind, scaled))
    model = SmallElasticPlastic(P)
    P.set_active_values_from_flat(offset, False)
    x = P.flat_active_values(True)
    obj = MPJVPObjective(Calibration(model, data, weight), F, make_newton_solve(model._residual))
    J, g = obj.evaluate_objective_and_grad(x)
    H = obj.evaluate_hessian(x)
    return dict(J=float(np.asarray(J)), grad=np.asarray(g, float), hessian=np.asarray(H, float),
                F=F, data=data, weight=weight, x_canonical=x, active_native=offset,
                active_idx=np.asarray(P.active_idx))


# --------------------------------------------------------------------------- #
#  D. FE element kernels: per_element_R_and_K_coupled / per_element_R_coupled  #
#     over the per-IP COUPLED evaluator (displacement and mixed u-p)            #
# --------------------------------------------------------------------------- #
def _fe_job(job):
    from cmad.fem.assembly import per_element_R_and_K_coupled, per_element_R_coupled
    from cmad.fem.element_family import ElementFamily
    from cmad.fem.interpolants import hex_linear, tet_linear
    from cmad.fem.mesh import _LOCAL_EDGES_PER_ELEMENT
    from cmad.fem.precompute import BlockIPGeometryPerElem, BlockIPGeometryShared
    from cmad.fem.quadrature import hex_quadrature, tet_quadrature
    from cmad.global_residuals.modes import GlobalResidualMode
    from cmad.global_residuals.small_disp_equilibrium import SmallDispEquilibrium
    from jax.flatten_util import ravel_pytree

    family, kind, mixed, seed, n_elems = job[:5]
    rate = len(job) > 5 and job[5] == "rate"
    rng = np.random.default_rng(seed)
    values = material(kind)
    P = parameters(values)
    if rate:
        from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic
        model = SmallRateElasticPlastic(P)
    else:
        model = SmallElasticPlastic(P)
    gr = SmallDispEquilibrium(ndims=3, mixed=mixed, stabilization_multiplier=1.0)
    ev = gr.for_model(model, GlobalResidualMode.COUPLED)            # local Newton 20 / 1e-12 / 1e-12
    unravel_xi = ravel_pytree(model._init_xi)[1]
    if family == "hex8":
        rule, interp, fam = hex_quadrature(2), hex_linear, ElementFamily.HEX_LINEAR
        Xref = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                         [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float) * 0.5
    else:
        rule, interp, fam = tet_quadrature(1), tet_linear, ElementFamily.TET_LINEAR
        Xref = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    n_b, n_ip = Xref.shape[0], len(rule.w)
    sh = [interp(np.asarray(x)) for x in rule.xi]
    N = np.array([np.asarray(s_.N) for s_ in sh])                    # (n_ip, n_b)
    gNref = np.array([np.asarray(s_.grad_N) for s_ in sh])           # (n_ip, n_b, 3)
    edges = np.asarray(_LOCAL_EDGES_PER_ELEMENT[fam])
    block_shapes = [(n_b, 3), (n_b, 1)] if mixed else [(n_b, 3)]
    shared = BlockIPGeometryShared(quad_w=np.asarray(rule.w),
                                   field_N_per_block=tuple(N for _ in block_shapes))
    rec = {k: [] for k in ("X", "U", "U_prev", "p", "xi_prev", "grad_N", "det", "h", "xi", "R_u", "R_p",
                           "K_uu", "K_up", "K_pu", "K_pp", "R_only_u", "R_only_p")}
    for e in range(n_elems):
        X = Xref + rng.normal(size=Xref.shape) * 0.06               # distorted element
        iso = np.einsum("ai,paj->pij", X, gNref)                     # precompute.py:239-241
        det = np.linalg.det(iso)
        gN = np.einsum("pnj,pji->pni", gNref, np.linalg.inv(iso))     # precompute.py:250-257
        ev_ = X[edges]
        h = float(np.sqrt(np.mean(np.sum((ev_[:, 1] - ev_[:, 0]) ** 2, axis=-1))))   # mesh.py:631-636
        geom = BlockIPGeometryPerElem(iso_jac_det=det, coords_ip=np.einsum("pa,ai->pi", N, X),
                                      field_grad_N_phys_per_block=tuple(gN for _ in block_shapes),
                                      element_size=h)
        xi_prev = np.zeros((n_ip, 7))
        U_last = np.zeros_like(X)
        for step in range(2):
            ramp = np.array([0.004, -0.001, 0.0015]) if step == 0 else np.array([0.003, 0.004, -0.002])
            U = X * ramp[None, :] * (1.0 + step) + rng.normal(size=X.shape) * 4e-4
            Ue = [U]
            if mixed:
                pe = rng.normal(size=(n_b, 1)) * 40.0 - 100.0
                Ue = [U, pe]
            Uprev = [np.zeros_like(u) for u in Ue]
            if rate:                   # the rate form sees eps(U) - eps(U_prev): carry the real previous step
                Uprev = [U_last.copy()] + ([np.zeros((n_b, 1))] if mixed else [])
            rec["U_prev"].append(Uprev[0].copy())
            Rb, Kb, xi = per_element_R_and_K_coupled(
                Ue, Uprev, P.values, xi_prev, geom, shared, ev["R_and_dR_dU_and_xi"], unravel_xi,
                {}, block_shapes, 0.0)
            Ronly = per_element_R_coupled(
                Ue, Uprev, P.values, xi_prev, geom, shared, ev["R"], unravel_xi, {}, block_shapes, 0.0)
            rec["X"].append(X); rec["U"].append(U); rec["xi_prev"].append(xi_prev.copy())
            rec["grad_N"].append(gN); rec["det"].append(det); rec["h"].append(h)
            rec["xi"].append(np.asarray(xi))
            rec["R_u"].append(np.asarray(Rb[0])); rec["R_only_u"].append(np.asarray(Ronly[0]))
            rec["K_uu"].append(np.asarray(Kb[0][0]).reshape(3 * n_b, 3 * n_b))
            if mixed:
                rec["p"].append(pe[:, 0])
                rec["R_p"].append(np.asarray(Rb[1])[:, 0]); rec["R_only_p"].append(np.asarray(Ronly[1])[:, 0])
                rec["K_up"].append(np.asarray(Kb[0][1]).reshape(3 * n_b, n_b))
                rec["K_pu"].append(np.asarray(Kb[1][0]).reshape(n_b, 3 * n_b))
                rec["K_pp"].append(np.asarray(Kb[1][1]).reshape(n_b, n_b))
            xi_prev = np.asarray(xi)
            U_last = U
    out = {k: np.array(v) for k, v in rec.items() if v}
    out["quad_w"], out["N"] = np.asarray(rule.w), N
    return out



# --------------------------------------------------------------------------- #
#  E. PLANE_STRESS / UNIAXIAL_STRESS deformation types (n_xi = 8 / 9)          #
# --------------------------------------------------------------------------- #
def deftype_F(def_type_name, nsteps=40):
    """Two-leg histories: plane stress as tests/objectives/test_J2_fd_checks.py:266-289
    (e_xx ramps to 0.02 then holds while e_yy ramps) plus a small in-plane shear;
    uniaxial stress: axial ramp to 0.02, partial unloading, reloading to 0.03."""
    h = nsteps // 2
    if def_type_name == "PLANE_STRESS":
        exx = np.r_[0.0, np.linspace(0.02 / h, 0.02, h), np.full(h, 0.02)]
        eyy = np.r_[0.0, np.zeros(h), np.linspace(0.02 / h, 0.02, h)]
        F = np.repeat(np.eye(2)[:, :, None], nsteps + 1, axis=2)
        F[0, 0] += exx; F[1, 1] += eyy
        F[0, 1] += 0.15 * exx; F[1, 0] += 0.05 * eyy          # unsymmetric grad_u: only sym part matters
        return F
    e = np.r_[0.0, np.linspace(0.02 / h, 0.02, h), np.linspace(0.02, 0.016, h // 2),
              np.linspace(0.016, 0.03, nsteps - h - h // 2)]
    return (1.0 + e).reshape(1, 1, nsteps + 1)


def _deftype_job(job):
    from cmad.models.deformation_types import DefType
    kind, dt_name = job[:2]
    rate = len(job) > 2 and job[2] == "rate"
    dt = DefType[dt_name]
    values = material(kind)
    P = parameters(values)
    if rate:          # the rate form under the def-types (small_rate_elastic_plastic.py:34-77, 296-345): n_xi 8 / 12
        from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic
        model = SmallRateElasticPlastic(P, def_type=dt)
    else:
        model = SmallElasticPlastic(P, def_type=dt)
    nxi = model.num_dofs
    F = deftype_F(dt_name)
    N = F.shape[2] - 1
    nd = F.shape[0]
    solve = make_newton_solve(model._residual)
    rec = {k: [] for k in ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp",
                           "traced_xi", "traced_iters", "dxi_dgradu", "dsig_dgradu")}
    model.set_xi_to_init_vals()
    for step in range(1, N + 1):
        U, Up = mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1])
        model.gather_global(U, Up)
        xp = [np.asarray(b).copy() for b in model.xi_prev()]
        # traced solver + IFT tangent from the same previous state
        xt = solve(xp, P.values, U, Up)
        rec["traced_iters"].append(int(_core.WHILE_LOG[-1][1][0]))
        rec["traced_xi"].append(np.concatenate([np.asarray(b) for b in xt]))

        def state_and_stress(U_):
            x = solve(xp, P.values, U_, Up)
            return jax.numpy.concatenate([jax.numpy.ravel(b) for b in x]), model.cauchy(x, xp, P.values, U_, Up)
        dx_dU, ds_dU = jax.jacfwd(state_and_stress)(U)
        rec["dxi_dgradu"].append(np.asarray(dx_dU.grad_fields["u"]).reshape(nxi, nd * nd))
        rec["dsig_dgradu"].append(np.asarray(ds_dU.grad_fields["u"]).reshape(9, nd * nd))
        # imperative solver on the Model object
        ii, cn = newton_solve(model)
        rec["iters"].append(ii); rec["cnorm"].append(cn)
        xi = [np.asarray(b).copy() for b in model.xi()]
        rec["xi"].append(np.concatenate(xi))
        model.seed_none(); model.evaluate_cauchy()
        rec["sigma"].append(vec6(model.Sigma()))
        jac = model._jacobian
        rec["dC_dxi"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI](xi, xp, P.values, U, Up)]))
        rec["dC_dxi_prev"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI_PREV](xi, xp, P.values, U, Up)]))
        dcdp = jac[DerivType.DPARAMS](xi, xp, P.values, U, Up)
        rec["dC_dp"].append(np.hstack([np.asarray(x).reshape(nxi, -1) for x in jax.tree_util.tree_leaves(dcdp)]))
        model.advance_xi()
    out = {k: np.array(v) for k, v in rec.items()}
    out["F"] = F
    # objectives (KA5): Calibration on the in-plane / axial stresses, offset parameters
    for scaled in (True, False):
        vals, act, tr = objective_trees(kind, scaled)
        Po = Parameters(vals, act, tr)
        mo_ = SmallRateElasticPlastic(Po, def_type=dt) if rate else SmallElasticPlastic(Po, def_type=dt)
        data = np.zeros((3, 3, N + 1))
        mo_.set_xi_to_init_vals()
        for step in range(1, N + 1):
            mo_.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
            newton_solve(mo_)
            mo_.seed_none(); mo_.evaluate_cauchy()
            data[:, :, step] = mo_.Sigma().copy()
            mo_.advance_xi()
        w = np.zeros((3, 3)); w[0, 0] = 1.0
        if dt_name == "PLANE_STRESS":
            w[1, 1] = 1.0; w[0, 1] = w[1, 0] = 0.5
        qoi = Calibration(mo_, data, w)
        offset = 1.1 * Po.flat_active_values(False)
        Po.set_active_values_from_flat(offset, False)
        x = Po.flat_active_values(True)
        tag = "scaled" if scaled else "native"
        for name, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
            Po.set_active_values_from_flat(offset, False)
            # MPObjective.__init__ (mp_objective.py:46) stores the model's CURRENT xi as the step-0 state
            # of the adjoint pass; after the data run above that is the END state.  Harmless while the
            # first load step is elastic (the identity-axes fixtures), wrong when it is plastic (the
            # rotated ones: adjoint != direct by 50 %): build the objective on a model at its initial state.
            mo_.set_xi_to_init_vals()
            J, g = ctor(qoi, F).evaluate(x)
            out[f"obj_{tag}.J_{name}"], out[f"obj_{tag}.grad_{name}"] = float(J), np.asarray(g, float)
        out[f"obj_{tag}.data"], out[f"obj_{tag}.weight"], out[f"obj_{tag}.x_canonical"] = data, w, x
        out[f"obj_{tag}.active_native"], out[f"obj_{tag}.active_idx"] = offset, np.asarray(Po.active_idx)
    return out



# --------------------------------------------------------------------------- #
#  F. SmallRateElasticPlastic (rate form), FULL_3D                              #
# --------------------------------------------------------------------------- #
def _rate_job(job):
    from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic
    kind, = job
    values = material(kind)
    P = parameters(values)
    model = SmallRateElasticPlastic(P)
    F = two_leg_F(7, 30, diag_only=kind == "hosford")
    N = F.shape[2] - 1
    solve = make_newton_solve(model._residual)
    rec = {k: [] for k in ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp",
                           "traced_xi", "traced_iters", "dxi_dgradu")}
    model.set_xi_to_init_vals()
    with np.errstate(all="ignore"):
        for step in range(1, N + 1):
            U, Up = mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1])
            model.gather_global(U, Up)
            xp = [np.asarray(b).copy() for b in model.xi_prev()]
            xt = solve(xp, P.values, U, Up)
            rec["traced_iters"].append(int(_core.WHILE_LOG[-1][1][0]))
            rec["traced_xi"].append(np.concatenate([np.asarray(b) for b in xt]))
            dx_dU = jax.jacfwd(lambda U_: jax.numpy.concatenate(
                [jax.numpy.ravel(b) for b in solve(xp, P.values, U_, Up)]))(U)
            rec["dxi_dgradu"].append(np.asarray(dx_dU.grad_fields["u"]).reshape(7, 9))
            ii, cn = newton_solve(model)
            rec["iters"].append(ii); rec["cnorm"].append(cn)
            xi = [np.asarray(b).copy() for b in model.xi()]
            rec["xi"].append(np.concatenate(xi))
            model.seed_none(); model.evaluate_cauchy()
            rec["sigma"].append(vec6(model.Sigma()))
            jac = model._jacobian
            rec["dC_dxi"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI](xi, xp, P.values, U, Up)]))
            rec["dC_dxi_prev"].append(np.hstack([np.asarray(b) for b in jac[DerivType.DXI_PREV](xi, xp, P.values, U, Up)]))
            dcdp = jac[DerivType.DPARAMS](xi, xp, P.values, U, Up)
            rec["dC_dp"].append(np.hstack([np.asarray(x).reshape(7, -1) for x in jax.tree_util.tree_leaves(dcdp)]))
            model.advance_xi()
    out = {k: np.array(v) for k, v in rec.items()}
    out["F"] = F
    return out



# --------------------------------------------------------------------------- #
#  G. UniaxialCalibration QoI (cmad/qois/uniaxial_calibration.py) on the       #
#     UNIAXIAL_STRESS deformation type: axial stress + the two off-axis strains #
# --------------------------------------------------------------------------- #
def _uniaxial_qoi_job(job):
    from cmad.models.deformation_types import DefType
    from cmad.qois.uniaxial_calibration import UniaxialCalibration
    kind, scaled = job[:2]
    F = deftype_F("UNIAXIAL_STRESS", nsteps=32)
    N = F.shape[2] - 1
    vals, act, tr = objective_trees(kind, scaled)
    Po = Parameters(vals, act, tr)
    if len(job) > 2 and job[2] == "rate":
        from cmad.models.small_rate_elastic_plastic import SmallRateElasticPlastic
        mo_ = SmallRateElasticPlastic(Po, def_type=DefType.UNIAXIAL_STRESS)
    else:
        mo_ = SmallElasticPlastic(Po, def_type=DefType.UNIAXIAL_STRESS)
    data = np.zeros((3, N + 1))
    mo_.set_xi_to_init_vals()
    for step in range(1, N + 1):
        mo_.gather_global(mp_U_from_F(F[:, :, step]), mp_U_from_F(F[:, :, step - 1]))
        newton_solve(mo_)
        mo_.seed_none(); mo_.evaluate_cauchy()
        xi = mo_.xi()
        data[:, step] = [mo_.Sigma()[0, 0], float(xi[2][0]) - 1.0, float(xi[2][1]) - 1.0]
        mo_.advance_xi()
    t = np.arange(N + 1)
    # per-step weights (weight_at_step = weight[:, step]): stress in its own units, strains scaled up
    weight = np.stack([1.0 + 0.25 * np.sin(0.7 * t), 2.0e4 * (1.0 + 0.1 * np.cos(0.3 * t)), 1.0e4 + 50.0 * t])
    qoi = UniaxialCalibration(mo_, data, weight, uniaxial_stress_idx=0, stretch_var_idx=2)
    offset = 1.1 * Po.flat_active_values(False)
    Po.set_active_values_from_flat(offset, False)
    x = Po.flat_active_values(True)
    out = {"F": F, "data": data, "weight": weight, "x_canonical": x, "active_native": offset,
           "active_idx": np.asarray(Po.active_idx)}
    for name, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        Po.set_active_values_from_flat(offset, False)
        mo_.set_xi_to_init_vals()             # see _objective_job: MPObjective stores the model's current xi as step 0
        J, g = ctor(qoi, F).evaluate(x)
        out[f"J_{name}"], out[f"grad_{name}"] = float(J), np.asarray(g, float)
    return out

# --------------------------------------------------------------------------- #
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    pool = mp.Pool(a.jobs)
    only = set(a.only.split(",")) if a.only else None

    if only is None or "traced" in only:
        out = {}
        for kind, key, n, steps, scale in (
                ("J2", "mp", 48, (12, 30, 48, 66, 84), 1.0),
                ("J2", "fe", 16, (20, 60, 90), 1.0),
                ("J2_kappa_mu", "mp", 8, (40, 80), 1.0),
                ("hill", "mp", 32, (15, 45, 75), 1.0),
                ("hill_rot", "fe", 32, (15, 45, 75), 1.0),
                ("hosford", "mp", 32, (15, 45, 75), 1.0),
                ("hosford_notch", "notch", 16, (25, 75), 2.0)):
            r = traced(pool, kind, key, n, steps, scale=scale)
            for k, v in r.items():
                out[f"{kind}.{key}.{k}"] = v
            print("traced", kind, key, "iters", np.bincount(r["iters"].ravel()), "flags", np.bincount(r["flags"].ravel()))
        np.savez_compressed(os.path.join(HERE, "ref_traced_newton.npz"), **out)

    if only is None or "imperative" in only:
        from oracle import analytic
        jobs, names = [], []
        for kind in ("J2", "hill", "hill_rot", "hosford"):
            for pname, F in (("uniaxial", None), ("biaxial", None),
                             ("twoleg", two_leg_F(7, 40, diag_only=kind == "hosford"))):
                if F is None:
                    mask = analytic.stress_masks_3d()[0 if pname == "uniaxial" else 1]
                    _, strain, _ = analytic.plastic_fields(mask, num_steps=30)
                    F = analytic.deformation_gradient_history(strain)
                jobs.append((kind, F)); names.append(f"{kind}.{pname}")
        out = {}
        for nm, r in zip(names, pool.map(_imperative_job, jobs)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("imperative", nm, "iters", np.bincount(r["iters"]))
        np.savez_compressed(os.path.join(HERE, "ref_imperative_newton.npz"), **out)

    if only is None or "objective" in only:
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("J2", "hill", "hosford"):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 24, scale=1.5, diag_only=kind == "hosford"), w))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_objective_job, jobs)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("objective", nm, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"])
        np.savez_compressed(os.path.join(HERE, "ref_mp_objectives.npz"), **out)

    if only is not None and "barlat" in only:
        # Yld2004-18p (effective_stress.py:81-84 -> verification/functions.py:71-154, jnp.linalg.eigh).
        # Not part of the default run: the nested duals through eigh make it slow (minutes per job).
        out = {}
        for kind, key, n, steps, tan in (
                ("barlat", "mp", 16, (15, 45, 75), False), ("barlat", "mp", 4, (30, 60), True),
                ("barlat_rot", "fe", 8, (20, 60), False), ("barlat_rot", "fe", 2, (20, 60), True),
                ("barlat_a8", "mp", 8, (20, 60), False)):
            r = traced(pool, kind, key, n, steps, seed=31 if tan else 22, chunk=1 if tan else 2, with_tangent=tan)
            for k, v in r.items():
                out[f"{kind}.{key}.{'tan.' if tan else ''}{k}"] = v
            print("traced", kind, key, tan, "iters", np.bincount(r["iters"].ravel()), "flags", np.bincount(r["flags"].ravel()), flush=True)
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs = [("barlat", sc, two_leg_F(11, 24, scale=1.5), w) for sc in (True, False)]
        for nm, r in zip(("barlat.scaled", "barlat.native"), pool.map(_objective_job, jobs)):
            for k, v in r.items():
                out[f"objective.{nm}.{k}"] = v
            print("objective", nm, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_barlat.npz"), **out)

    if only is None or "hessian" in only:
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("J2", "hill", "hosford"):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 12, scale=1.5, diag_only=kind == "hosford"), w))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]))
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian.npz"), **out)

    if only is None or "hessian_rot" in only:
        # the direct-adjoint Hessian with rotated material axes (anisotropic Hill, two hardening laws)
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("hill_rot",):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 10, scale=1.5), w))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_rot", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]))
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_rot.npz"), **out)

    if only is None or "hessian_rate" in only:
        # the direct-adjoint Hessian of the rate form (FULL_3D; identity and rotated material axes)
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("J2", "hill", "hosford", "hill_rot"):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 10, scale=1.5, diag_only=kind == "hosford"), w,
                             "FULL_3D", "rate"))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_rate", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_rate.npz"), **out)

    if only is None or "hessian_jvp" in only:
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("J2", "hill", "hosford"):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 8, scale=1.5, diag_only=kind == "hosford"), w))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_jvp_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_jvp", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]))
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_jvp.npz"), **out)

    if only is None or "hessian_dt" in only:
        # KA5 setting (tests/objectives/test_jvp_vs_original.py, test_J2_fd_checks.py:266-289): the
        # Hessian in PLANE_STRESS (and UNIAXIAL_STRESS), in-plane / axial stress data only
        jobs, names = [], []
        for kind, scaled, dt_name in (("J2", True, "PLANE_STRESS"), ("J2", False, "PLANE_STRESS"),
                                      ("hill", False, "PLANE_STRESS"), ("hosford", True, "PLANE_STRESS"),
                                      ("J2", True, "UNIAXIAL_STRESS"), ("J2", False, "UNIAXIAL_STRESS")):
            w = np.zeros((3, 3)); w[0, 0] = 1.0
            if dt_name == "PLANE_STRESS":
                w[1, 1] = 1.0; w[0, 1] = w[1, 0] = 0.5
            jobs.append((kind, scaled, deftype_F(dt_name, nsteps=12), w, dt_name))
            names.append(f"{kind}.{'scaled' if scaled else 'native'}.{dt_name}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            if not np.all(np.isfinite(r["hessian"])):
                # a non-finite entry would be an artefact of the NumPy stand-in, not a fixture
                print("hessian_dt", nm, "NON-FINITE Hessian - case dropped")
                continue
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_dt", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]))
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_dt.npz"), **out)

    if only is None or "hessian_rate_dt" in only:
        # the rate form's Hessian in PLANE_STRESS / UNIAXIAL_STRESS (n_xi 8 / 12), identity and rotated axes
        jobs, names = [], []
        for kind, scaled, dt_name in (("J2", True, "PLANE_STRESS"), ("J2", False, "PLANE_STRESS"),
                                      ("hill_rot", False, "PLANE_STRESS"), ("hosford", True, "PLANE_STRESS"),
                                      ("J2", True, "UNIAXIAL_STRESS"), ("hill_rot", False, "UNIAXIAL_STRESS")):
            w = np.zeros((3, 3)); w[0, 0] = 1.0
            if dt_name == "PLANE_STRESS":
                w[1, 1] = 1.0; w[0, 1] = w[1, 0] = 0.5
            jobs.append((kind, scaled, deftype_F(dt_name, nsteps=10), w, dt_name, "rate"))
            names.append(f"{kind}.{'scaled' if scaled else 'native'}.{dt_name}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            if not np.all(np.isfinite(r["hessian"])):
                print("hessian_rate_dt", nm, "NON-FINITE Hessian - case dropped")
                continue
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_rate_dt", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_rate_dt.npz"), **out)

    if only is None or "hessian_dt_rot" in only:
        # SmallElasticPlastic's Hessian in PLANE_STRESS / UNIAXIAL_STRESS with rotated material axes
        jobs, names = [], []
        for kind, scaled, dt_name in (("hill_rot", True, "PLANE_STRESS"), ("hill_rot", False, "PLANE_STRESS"),
                                      ("hill_rot", True, "UNIAXIAL_STRESS"), ("hill_rot", False, "UNIAXIAL_STRESS")):
            w = np.zeros((3, 3)); w[0, 0] = 1.0
            if dt_name == "PLANE_STRESS":
                w[1, 1] = 1.0; w[0, 1] = w[1, 0] = 0.5
            jobs.append((kind, scaled, deftype_F(dt_name, nsteps=10), w, dt_name))
            names.append(f"{kind}.{'scaled' if scaled else 'native'}.{dt_name}")
        out = {}
        for nm, r in zip(names, pool.map(_hessian_job, jobs, chunksize=1)):
            if not np.all(np.isfinite(r["hessian"])):
                print("hessian_dt_rot", nm, "NON-FINITE Hessian - case dropped")
                continue
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("hessian_dt_rot", nm, r["J"], r["grad"], np.linalg.eigvalsh(r["hessian"]), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_mp_hessian_dt_rot.npz"), **out)

    if only is None or "fe" in only:
        jobs, names = [], []
        for family in ("tet4", "hex8"):
            for kind in ("J2", "hill_rot", "hosford"):
                for mixed in (False, True):
                    jobs.append((family, kind, mixed, 5 + len(jobs), 3 if family == "tet4" else 2))
                    names.append(f"{family}.{kind}.{'mixed' if mixed else 'disp'}")
        out = {}
        for nm, r in zip(names, pool.map(_fe_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("fe", nm, "alpha max", r["xi"][..., 6].max())
        np.savez_compressed(os.path.join(HERE, "ref_fe_elements.npz"), **out)

    if only is not None and "partials" in only:
        jobs = [("J2", "mp", 6, (20, 60), 5), ("hill_rot", "fe", 6, (20, 60), 6), ("hosford", "mp", 6, (20, 60), 7),
                ("barlat_rot", "fe", 3, (20, 60), 8)]
        out = {}
        for job, r in zip(jobs, pool.map(_partials_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{job[0]}.{k}"] = v
            print("partials", job[0], "plastic", int((r["flags"] > 0).sum()), "of", r["flags"].size, flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_model_partials.npz"), **out)

    if only is not None and "barlat_fe" in only:
        # element blocks with the Yld2004-18p surface (slow: eigh under nested duals at every point)
        jobs = [("tet4", "barlat_rot", False, 41, 2), ("tet4", "barlat", True, 42, 2),
                ("hex8", "barlat", False, 43, 1), ("hex8", "barlat_rot", True, 44, 1)]
        names = [f"{f}.{k}.{'mixed' if m else 'disp'}" for f, k, m, _, _ in jobs]
        out = {}
        for nm, r in zip(names, pool.map(_fe_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("fe", nm, "alpha max", r["xi"][..., 6].max(), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_barlat_fe.npz"), **out)

    if only is None or "deftypes" in only:
        jobs = [(kind, dt) for dt in ("PLANE_STRESS", "UNIAXIAL_STRESS") for kind in ("J2", "hill", "hosford")]
        out = {}
        for (kind, dt), r in zip(jobs, pool.map(_deftype_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{dt}.{k}"] = v
            print("deftypes", kind, dt, "iters", np.bincount(r["iters"]), "traced", np.bincount(r["traced_iters"]),
                  "alpha", r["xi"][-1, 6], "J", r["obj_scaled.J_adjoint"])
        np.savez_compressed(os.path.join(HERE, "ref_def_types.npz"), **out)

    if only is None or "rate" in only:
        jobs = [(k,) for k in ("J2", "hill", "hosford")]
        out = {}
        for (kind,), r in zip(jobs, pool.map(_rate_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{k}"] = v
            print("rate", kind, "iters", np.bincount(r["iters"]), "traced", np.bincount(r["traced_iters"]),
                  "alpha", r["xi"][-1, 6])
        np.savez_compressed(os.path.join(HERE, "ref_rate_model.npz"), **out)

    if only is None or "uniaxial_qoi" in only:
        jobs = [(k, sc) for k in ("J2", "hill", "hosford") for sc in (True, False)]
        out = {}
        for (kind, sc), r in zip(jobs, pool.map(_uniaxial_qoi_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{'scaled' if sc else 'native'}.{k}"] = v
            print("uniaxial qoi", kind, sc, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"])
        np.savez_compressed(os.path.join(HERE, "ref_uniaxial_qoi.npz"), **out)

    if only is None or "uniaxial_qoi_rate" in only:
        # UniaxialCalibration on the rate form (stretch block 2 as well), identity and rotated axes
        jobs = [(k, sc, "rate") for k in ("J2", "hill_rot", "hosford") for sc in (True, False)]
        out = {}
        for (kind, sc, _), r in zip(jobs, pool.map(_uniaxial_qoi_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{'scaled' if sc else 'native'}.{k}"] = v
            print("uniaxial qoi rate", kind, sc, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_uniaxial_qoi_rate.npz"), **out)

    if only is None or "rate_objective" in only:
        # MPAdjointObjective / MPDirectObjective + Calibration over SmallRateElasticPlastic
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs, names = [], []
        for kind in ("J2", "hill", "hosford"):
            for scaled in (True, False):
                jobs.append((kind, scaled, two_leg_F(11, 24, scale=1.5, diag_only=kind == "hosford"), w, "rate"))
                names.append(f"{kind}.{'scaled' if scaled else 'native'}")
        out = {}
        for nm, r in zip(names, pool.map(_objective_job, jobs)):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("rate objective", nm, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"])
        np.savez_compressed(os.path.join(HERE, "ref_rate_objectives.npz"), **out)

    if only is None or "rate_fe" in only:
        # per_element_R_and_K_coupled over the rate form's per-IP COUPLED evaluator (displacement form)
        jobs = [(fam, kind, False, 31 + i, 3, "rate")
                for i, (fam, kind) in enumerate((("tet4", "J2"), ("hex8", "J2"), ("tet4", "hill"), ("hex8", "hosford")))]
        out = {}
        for job, r in zip(jobs, pool.map(_fe_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{job[0]}.{job[1]}.{k}"] = v
            print("rate fe", job[:2], "alpha max", r["xi"][..., 6].max(), "|K|", np.abs(r["K_uu"]).max())
        np.savez_compressed(os.path.join(HERE, "ref_rate_fe_elements.npz"), **out)


    if only is not None and "deftypes_rot" in only:
        # PLANE_STRESS / UNIAXIAL_STRESS with rotated material axes (small_elastic_plastic.py:44-62:
        # the uniaxial constraint is imposed on the GLOBAL off-diagonal strains, Q ep Q^T)
        jobs = [("hill_rot", dt) for dt in ("PLANE_STRESS", "UNIAXIAL_STRESS")]
        out = {}
        for (kind, dt), r in zip(jobs, pool.map(_deftype_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{dt}.{k}"] = v
            print("deftypes_rot", kind, dt, "iters", np.bincount(r["iters"]), "traced", np.bincount(r["traced_iters"]),
                  "alpha", r["xi"][-1, 6], "J", r["obj_scaled.J_adjoint"], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_def_types_rot.npz"), **out)

    if only is not None and "barlat_more" in only:
        # Yld2004-18p in the def-type kernels and in the rate form (material point): slow jobs, own files
        dt_jobs = [("barlat", dt) for dt in ("PLANE_STRESS", "UNIAXIAL_STRESS")]
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        ob_jobs = [("barlat", sc, two_leg_F(11, 24, scale=1.5), w, "rate") for sc in (True, False)]
        r_dt = pool.map_async(_deftype_job, dt_jobs, chunksize=1)
        r_rm = pool.map_async(_rate_job, [("barlat",)], chunksize=1)
        r_ob = pool.map_async(_objective_job, ob_jobs, chunksize=1)
        out = {}
        for (kind, dt), r in zip(dt_jobs, r_dt.get()):
            for k, v in r.items():
                out[f"{kind}.{dt}.{k}"] = v
            print("barlat deftypes", dt, "iters", np.bincount(r["iters"]), "alpha", r["xi"][-1, 6], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_def_types_barlat.npz"), **out)
        out = {}
        (r,) = r_rm.get()
        for k, v in r.items():
            out[f"barlat.{k}"] = v
        print("barlat rate iters", np.bincount(r["iters"]), "alpha", r["xi"][-1, 6], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_rate_model_barlat.npz"), **out)
        out = {}
        for nm, r in zip(("barlat.scaled", "barlat.native"), r_ob.get()):
            for k, v in r.items():
                out[f"{nm}.{k}"] = v
            print("barlat rate objective", nm, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_rate_objectives_barlat.npz"), **out)

    if only is not None and "rate_deftypes" in only:
        jobs = [(kind, dt, "rate") for dt in ("PLANE_STRESS", "UNIAXIAL_STRESS") for kind in ("J2", "hill_rot", "hosford")]
        out = {}
        for (kind, dt, _), r in zip(jobs, pool.map(_deftype_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{dt}.{k}"] = v
            print("rate deftypes", kind, dt, "iters", np.bincount(r["iters"]), "traced", np.bincount(r["traced_iters"]),
                  "alpha", r["xi"][-1, 6], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_def_types_rate.npz"), **out)

    if only is not None and "rate_deftypes_barlat" in only:
        # Yld2004-18p in the rate form under the def-types (slow jobs, own file)
        jobs = [("barlat", dt, "rate") for dt in ("PLANE_STRESS", "UNIAXIAL_STRESS")]
        out = {}
        for (kind, dt, _), r in zip(jobs, pool.map(_deftype_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"{kind}.{dt}.{k}"] = v
            print("rate deftypes", kind, dt, "iters", np.bincount(r["iters"]), "traced", np.bincount(r["traced_iters"]),
                  "alpha", r["xi"][-1, 6], flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_def_types_rate_barlat.npz"), **out)

    if only is not None and "rate_rot" in only:
        # SmallRateElasticPlastic with rotated material axes (the case tests/models/
        # test_hill_material_rotations.py runs) and in the mixed u-p formulation (tests/fem/
        # test_mixed_up_plastic.py): K1 / Model AD products, calibration objectives, element blocks
        out = {}
        (r,) = pool.map(_rate_job, [("hill_rot",)], chunksize=1)
        for k, v in r.items():
            out[f"model.hill_rot.{k}"] = v
        print("rate_rot model iters", np.bincount(r["iters"]), "alpha", r["xi"][-1, 6], flush=True)
        w = np.array([[1.0, 0.5, 0.0], [0.5, 1.0, 0.0], [0.0, 0.0, 0.25]])
        jobs = [("hill_rot", sc, two_leg_F(11, 24, scale=1.5), w, "rate") for sc in (True, False)]
        for nm, r in zip(("scaled", "native"), pool.map(_objective_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"objective.hill_rot.{nm}.{k}"] = v
            print("rate_rot objective", nm, r["J_adjoint"], r["grad_adjoint"], r["grad_direct"], flush=True)
        jobs = [("tet4", "hill_rot", False, 51, 3, "rate"), ("hex8", "hill_rot", True, 52, 2, "rate"),
                ("tet4", "J2", True, 53, 3, "rate"), ("hex8", "hosford", True, 54, 2, "rate")]
        for job, r in zip(jobs, pool.map(_fe_job, jobs, chunksize=1)):
            for k, v in r.items():
                out[f"fe.{job[0]}.{job[1]}.{'mixed' if job[2] else 'disp'}.{k}"] = v
            print("rate_rot fe", job[:3], "alpha max", r["xi"][..., 6].max(), "|K|", np.abs(r["K_uu"]).max(), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_rate_rot.npz"), **out)


if __name__ == "__main__":
    main()
