"""FE element kernels against golden vectors produced by EXECUTING THE
REFERENCE'S OWN SOURCE (tests/golden/make_reference_golden.py, section D):
`per_element_R_and_K_coupled` / `per_element_R_coupled` (cmad/fem/assembly.py:416-613)
over the per-IP COUPLED evaluator of `SmallDispEquilibrium.for_model`
(global_residual.py:341-400, small_disp_equilibrium.py:82-118), displacement AND
mixed u-p formulations, tet4 (1 IP) and hex8 (8 IPs), J2 / rotated anisotropic
Hill with two hardening laws / Hosford, on distorted elements, two load steps
(the second from a non-zero plastic history).  Fixture: tests/golden/ref_fe_elements.npz.

CPU: the NumPy/C++ FE oracle (`oracle/fe_oracle.py`) reproduces R_e, every K block
and xi; plus a finite-difference check of the mixed oracle's tangent blocks.
GPU: the CUDA element-block kernels through the C-ABI do the same.
"""
import os

import numpy as np
import pytest

from oracle import fe_oracle, oracle_c as oc
from tests.golden.materials import material
from tests.helpers import rel_err

FE = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fe_elements.npz"))
CASES = sorted({k.rsplit(".", 1)[0] for k in FE.files})
NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)          # for_model's COUPLED defaults


def _case_arrays(case, FE=FE):
    """Every (element, step) record of the fixture as one independent element of a block
    with private nodes: dofs block-major (u: 3*n_nodes, then p)."""
    family, kind, form = case.split(".")
    g = {k: FE[f"{case}.{k}"] for k in ("U", "xi_prev", "grad_N", "det", "h", "xi", "R_u", "K_uu", "R_only_u")}
    n_e, n_b = g["U"].shape[0], g["U"].shape[1]
    n_nodes = n_e * n_b
    conn = np.arange(n_nodes).reshape(n_e, n_b)
    eq_u = (conn[:, :, None] * 3 + np.arange(3)[None, None, :]).reshape(n_e, 3 * n_b)
    U = g["U"].reshape(-1)
    mixed = form == "mixed"
    eq_p = None
    if mixed:
        for k in ("p", "R_p", "K_up", "K_pu", "K_pp", "R_only_p"):
            g[k] = FE[f"{case}.{k}"]
        eq_p = 3 * n_nodes + conn
        U = np.concatenate([U, g["p"].reshape(-1)])
    return family, kind, mixed, g, eq_u, eq_p, U, FE[f"{case}.quad_w"], FE[f"{case}.N"]


def _compare(case, out, g, mixed, tol=1e-9):
    assert rel_err(out["xi"], g["xi"]) < 1e-10, (case, "xi")
    assert rel_err(out["R_u"], g["R_u"].reshape(out["R_u"].shape)) < tol, (case, "R_u")
    assert rel_err(out["K_uu"], g["K_uu"]) < tol, (case, "K_uu", rel_err(out["K_uu"], g["K_uu"]))
    if mixed:
        for k in ("R_p", "K_up", "K_pu", "K_pp"):
            assert rel_err(out[k], g[k]) < tol, (case, k, rel_err(out[k], g[k]))
    # the residual-only evaluator of the reference agrees with its fused one
    assert rel_err(g["R_only_u"], g["R_u"]) < 1e-12


@pytest.mark.parametrize("case", CASES)
def test_fe_oracle_vs_reference_elements(case):
    family, kind, mixed, g, eq_u, eq_p, U, quad_w, N = _case_arrays(case)
    prob = oc.describe(material(kind), None, newton_mode="traced", strain_comps=9, **NEWTON)
    assert g["xi"][..., 6].max() > 0, "fixture must be plastic"
    if mixed:
        out = fe_oracle.assemble_block_mixed(prob, eq_u, eq_p, U, g["xi_prev"], g["grad_N"], N, g["det"],
                                             quad_w, g["h"], stab_mult=1.0)
    else:
        r = fe_oracle.assemble_block(prob, eq_u, U, g["xi_prev"], g["grad_N"], g["det"], quad_w)
        out = {"xi": r["xi"], "R_u": r["R_elem"], "K_uu": r["K_elem"]}
    _compare(case, out, g, mixed)


@pytest.mark.parametrize("family", ["tet4", "hex8"])
def test_mixed_oracle_tangent_blocks_vs_finite_differences(family):
    """d[R_u, R_p]/d[U, p] of the Newton-running mixed block by central differences."""
    case = f"{family}.J2.mixed"
    _, kind, mixed, g, eq_u, eq_p, U, quad_w, N = _case_arrays(case)
    prob = oc.describe(material(kind), None, newton_mode="traced", strain_comps=9, **NEWTON)
    sl = slice(0, 1)                                                   # first element record

    def run(Uv, want_K):
        return fe_oracle.assemble_block_mixed(prob, eq_u[sl], eq_p[sl], Uv, g["xi_prev"][sl], g["grad_N"][sl], N,
                                              g["det"][sl], quad_w, g["h"][sl], want_K=want_K)
    base = run(U, True)
    n_b = eq_p.shape[1]
    cols_u, cols_p = eq_u[0], eq_p[0]
    Kfd_u = np.zeros((4 * n_b, 3 * n_b)); Kfd_p = np.zeros((4 * n_b, n_b))
    for cols, Kfd, eps in ((cols_u, Kfd_u, 1e-7), (cols_p, Kfd_p, 1e-3)):
        for j, c in enumerate(cols):
            Up, Um = U.copy(), U.copy()
            Up[c] += eps; Um[c] -= eps
            rp, rm = run(Up, False), run(Um, False)
            Kfd[:, j] = np.concatenate([(rp["R_u"] - rm["R_u"])[0], (rp["R_p"] - rm["R_p"])[0]]) / (2 * eps)
    K_an_u = np.vstack([base["K_uu"][0], base["K_pu"][0]])
    K_an_p = np.vstack([base["K_up"][0], base["K_pp"][0]])
    assert rel_err(Kfd_u, K_an_u) < 2e-6
    assert rel_err(Kfd_p, K_an_p) < 1e-8


# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("deterministic", [True, False])
def test_cuda_fe_vs_reference_elements(cuda_device, case, deterministic):
    _cuda_fe_vs_fixture(cuda_device, case, deterministic, FE)


def _cuda_fe_vs_fixture(cuda_device, case, deterministic, FE):
    import torch
    from cmad_b200 import fe, material_from_values
    from cmad_b200.fe_mesh import FEBlockArrays
    family, kind, mixed, g, eq_u, eq_p, U, quad_w, N = _case_arrays(case, FE)
    n_e, n_b = eq_u.shape[0], eq_u.shape[1] // 3
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(cuda_device)  # noqa: E731
    arr = FEBlockArrays(t(eq_u, torch.int32), t(g["grad_N"]), t(g["det"]), t(quad_w), t(N), int(U.size),
                        t(eq_p, torch.int32) if mixed else None, t(g["h"]) if mixed else None)
    mat = material_from_values(material(kind))
    nw = fe.fe_newton_settings(**NEWTON)
    Ut, xp = t(U), t(g["xi_prev"])
    if mixed:
        plan = fe.mixed_r_plan(arr) if deterministic else None
        R, vals, xi = fe.assemble_element_block_mixed(mat, nw, arr, Ut, xp, stab_mult=1.0, r_plan=plan)
        torch.cuda.synchronize()
        v = vals.cpu().numpy()
        nu, npd = 3 * n_b, n_b
        sizes = np.cumsum([0, n_e * nu * nu, n_e * nu * npd, n_e * npd * nu, n_e * npd * npd])
        Rn = R.cpu().numpy()
        out = {"xi": xi.cpu().numpy(), "R_u": Rn[eq_u], "R_p": Rn[eq_p],      # private nodes: scatter is a permutation
               "K_uu": v[sizes[0]:sizes[1]].reshape(n_e, nu, nu), "K_up": v[sizes[1]:sizes[2]].reshape(n_e, nu, npd),
               "K_pu": v[sizes[2]:sizes[3]].reshape(n_e, npd, nu), "K_pp": v[sizes[3]:sizes[4]].reshape(n_e, npd, npd)}
        # residual-only variant (K4)
        R2, none, _ = fe.assemble_element_block_mixed(mat, nw, arr, Ut, xp, stab_mult=1.0, r_plan=plan, want_K=False)
        assert none is None and rel_err(R2.cpu().numpy(), Rn) < 1e-13
    else:
        plan = fe.SegmentPlan(eq_u.reshape(-1), int(U.size), device=cuda_device) if deterministic else None
        R, vals, xi = fe.assemble_element_block(mat, nw, arr, Ut, xp, r_plan=plan)
        torch.cuda.synchronize()
        out = {"xi": xi.cpu().numpy(), "R_u": R.cpu().numpy()[eq_u],
               "K_uu": vals.cpu().numpy().reshape(n_e, 3 * n_b, 3 * n_b)}
    _compare(case, out, g, mixed)
