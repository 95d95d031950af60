"""SmallRateElasticPlastic with ROTATED MATERIAL AXES (the configuration the reference's
tests/models/test_hill_material_rotations.py runs on both small-strain models) and in the MIXED u-p
formulation (tests/fem/test_mixed_up_plastic.py::test_small_rate_elastic_plastic), against the
reference's own run - tests/golden/ref_rate_rot.npz, written by `make_reference_golden.py --only
rate_rot`: the `Model` object on an anisotropic Hill material with rotated axes (imperative and traced
Newton, AD products), `MPAdjointObjective` / `MPDirectObjective`, and `per_element_R_and_K_coupled`
over `SmallDispEquilibrium(mixed=...)` on distorted tet4 / hex8 elements.

In the mixed form of this model `hydro_cauchy = tr(cauchy(xi)) / 3`
(small_rate_elastic_plastic.py:369-376): the pressure rows depend on the local state, K_pu carries
the trace rows of the IFT tangent.  CPU: torch-AD oracle; GPU: the CUDA kernels through the C-ABI."""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import active_all_scalars, const_like, material, objective_trees
from tests.helpers import UP, rel_err

RR = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rate_rot.npz"))
FE_CASES = sorted({k.split(".", 1)[1].rsplit(".", 1)[0] for k in RR.files if k.startswith("fe.")})
FE_NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)


def test_fixture_contents():
    assert FE_CASES == ["hex8.hill_rot.mixed", "hex8.hosford.mixed", "tet4.J2.mixed", "tet4.hill_rot.disp"]
    assert RR["model.hill_rot.xi"][-1, 6] > 0
    for case in FE_CASES:
        assert RR[f"fe.{case}.xi"][..., 6].max() > 0
    assert rel_err(RR["objective.hill_rot.native.grad_direct"], RR["objective.hill_rot.native.grad_adjoint"]) < 1e-9


def test_torch_oracle_vs_reference_rotated_rate_model():
    import torch
    from oracle import cmad_oracle as co
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    tv = co.to_torch_tree(material("hill_rot"))
    F = RR["model.hill_rot.F"]
    for t in (3, 12, 20, 29):
        gu, gup = torch.from_numpy(F[:, :, t] - np.eye(3)), torch.from_numpy(F[:, :, t - 1] - np.eye(3))
        xp = RR["model.hill_rot.xi"][t - 2] if t > 1 else np.zeros(7)
        x, info = co.newton_imperative(xp.copy(), tv, gu, gup, spec)
        assert info.iters == RR["model.hill_rot.iters"][t - 1]
        assert rel_err(x.numpy(), RR["model.hill_rot.xi"][t - 1]) < 1e-10
        sig = co.rate_cauchy(x, torch.from_numpy(xp.copy()), tv, gu, gup, spec).numpy()
        assert rel_err(np.array([sig[i, j] for i, j in UP]), RR["model.hill_rot.sigma"][t - 1]) < 1e-10


@pytest.mark.parametrize("case", ["tet4.J2.mixed", "hex8.hill_rot.mixed"])
def test_torch_oracle_mixed_rate_elements_vs_reference(case):
    import torch
    from oracle import cmad_oracle as co
    family, kind, form = case.split(".")
    g = {k: RR[f"fe.{case}.{k}"] for k in ("U", "U_prev", "p", "xi_prev", "grad_N", "det", "h", "xi", "R_u", "R_p",
                                           "K_uu", "K_up", "K_pu", "K_pp")}
    quad_w, N = RR[f"fe.{case}.quad_w"], RR[f"fe.{case}.N"]
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    tv = co.to_torch_tree(material(kind))
    for e in (1, g["U"].shape[0] - 1):                      # second steps: a plastic previous state
        n_b = g["U"].shape[1]
        acc = [0.0] * 6
        for ip in range(len(quad_w)):
            r = co.coupled_ip_mixed(tv, g["U"][e], g["p"][e], g["U_prev"][e], g["xi_prev"][e, ip], g["grad_N"][e, ip],
                                    N[ip], float(quad_w[ip]), float(g["det"][e, ip]), float(g["h"][e]), spec,
                                    newton_settings=FE_NEWTON)
            acc = [a + np.asarray(v) for a, v in zip(acc, r[:6])]
            assert rel_err(r[6].numpy(), g["xi"][e, ip]) < 1e-10
        for got, key in zip(acc, ("R_u", "R_p", "K_uu", "K_up", "K_pu", "K_pp")):
            assert rel_err(got.reshape(g[key][e].shape), g[key][e]) < 1e-8, (case, e, key)


# ------------------------------------------------------------------------------------------ #
#  GPU                                                                                        #
# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
def test_cuda_vs_reference_rotated_rate_model(cuda_device):
    """K1-rate with rotated axes: imperative and traced Newton, AD products, IFT tangents."""
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    kind = "hill_rot"
    values = material(kind)
    P = Parameters(values, active_all_scalars(values), const_like(values, None))
    mat = material_from_values(values, model="small_rate_elastic_plastic")
    pid, aidx = active_param_ids(P), np.asarray(P.active_idx)
    F = RR[f"model.{kind}.F"]
    want = ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp", "dxi_deps", "dsig_deps")
    cols = [((3 * k + l,) if k == l else (3 * k + l, 3 * l + k)) for k, l in UP]
    xi = torch.zeros((7, 1), dtype=torch.float64, device=cuda_device)
    g = lambda k: RR[f"model.{kind}.{k}"]  # noqa: E731
    for t in range(1, F.shape[2]):
        e = torch.from_numpy((F[:, :, t] - F[:, :, t - 1]).reshape(9, 1).copy()).to(cuda_device)
        ot = mp.mp_update(mat, NewtonSettings(mode="traced"), pid, xi, e, outputs=want)
        assert int(ot["iters"][0]) == g("traced_iters")[t - 1], (t, "traced count")
        assert rel_err(ot["xi"][:, 0].cpu().numpy(), g("traced_xi")[t - 1]) < 1e-10
        dx = np.stack([sum(g("dxi_dgradu")[t - 1][:, c] for c in cc) for cc in cols], axis=-1)   # (7, 6)
        assert rel_err(ot["dxi_deps"][:, 0].cpu().numpy().reshape(7, 6), dx) < 1e-8, (t, "dxi_deps")
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), pid, xi, e, outputs=want)
        assert int(o["iters"][0]) == g("iters")[t - 1], t
        assert abs(float(o["cnorm"][0]) - g("cnorm")[t - 1]) < 1e-11
        assert rel_err(o["xi"][:, 0].cpu().numpy(), g("xi")[t - 1]) < 1e-10, t
        assert rel_err(o["sigma"][:, 0].cpu().numpy(), g("sigma")[t - 1]) < 1e-10, (t, "global cauchy")
        assert rel_err(o["dC_dxi"][:, 0].cpu().numpy().reshape(7, 7), g("dC_dxi")[t - 1]) < 1e-9
        assert rel_err(o["dC_dxi_prev"][:, 0].cpu().numpy().reshape(7, 7), g("dC_dxi_prev")[t - 1]) < 1e-9
        assert rel_err(o["dC_dp"][:, 0].cpu().numpy().reshape(7, len(aidx)), g("dC_dp")[t - 1][:, aidx]) < 1e-9
        xi = o["xi"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["scaled", "native"])
def test_cuda_rotated_rate_objectives_vs_reference(cuda_device, mode):
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectObjective, SmallRateElasticPlastic
    case = f"objective.hill_rot.{mode}"
    values, act, tr = objective_trees("hill_rot", mode == "scaled")
    for strategy, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        P = Parameters(values, act, tr)
        assert np.array_equal(P.active_idx, RR[f"{case}.active_idx"])
        obj = ctor(Calibration(SmallRateElasticPlastic(P), RR[f"{case}.data"], RR[f"{case}.weight"]),
                   RR[f"{case}.F"], device=cuda_device)
        r = obj.evaluate(RR[f"{case}.x_canonical"])
        assert abs(r.J - RR[f"{case}.J_{strategy}"]) < 1e-11 * abs(r.J), (case, strategy)
        assert rel_err(r.grad, RR[f"{case}.grad_{strategy}"]) < 1e-9, (case, strategy, r.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("case", FE_CASES)
def test_cuda_rate_fe_blocks_rotated_and_mixed_vs_reference(cuda_device, case):
    import torch
    from cmad_b200 import fe, material_from_values
    from cmad_b200.fe_mesh import FEBlockArrays
    family, kind, form = case.split(".")
    mixed = form == "mixed"
    keys = ("U", "U_prev", "xi_prev", "grad_N", "det", "h", "xi", "R_u", "K_uu") + \
        (("p", "R_p", "K_up", "K_pu", "K_pp") if mixed else ())
    g = {k: RR[f"fe.{case}.{k}"] for k in keys}
    quad_w, N = RR[f"fe.{case}.quad_w"], RR[f"fe.{case}.N"]
    n_e, n_b = g["U"].shape[0], g["U"].shape[1]
    n_nodes = n_e * n_b
    conn = np.arange(n_nodes).reshape(n_e, n_b)
    eq_u = (conn[:, :, None] * 3 + np.arange(3)[None, None, :]).reshape(n_e, 3 * n_b)
    eq_p = 3 * n_nodes + conn
    U = np.concatenate([g["U"].reshape(-1), g["p"].reshape(-1)]) if mixed else g["U"].reshape(-1)
    Up = np.concatenate([g["U_prev"].reshape(-1), np.zeros(n_nodes)]) if mixed else g["U_prev"].reshape(-1)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(cuda_device)  # noqa: E731
    arr = FEBlockArrays(t(eq_u, torch.int32), t(g["grad_N"]), t(g["det"]), t(quad_w), t(N), int(U.size),
                        t(eq_p, torch.int32) if mixed else None, t(g["h"]) if mixed else None)
    mat = material_from_values(material(kind), model="small_rate_elastic_plastic")
    nw = fe.fe_newton_settings(**FE_NEWTON)
    Ut, Upt, xp = t(U), t(Up), t(g["xi_prev"])
    if not mixed:
        plan = fe.SegmentPlan(eq_u.reshape(-1), int(U.size), device=cuda_device)
        R, vals, xi = fe.assemble_element_block(mat, nw, arr, Ut, xp, r_plan=plan, U_prev=Upt)
        torch.cuda.synchronize()
        assert rel_err(xi.cpu().numpy(), g["xi"]) < 1e-10
        assert rel_err(R.cpu().numpy()[eq_u], g["R_u"].reshape(n_e, -1)) < 1e-9
        assert rel_err(vals.cpu().numpy().reshape(n_e, 3 * n_b, 3 * n_b), g["K_uu"]) < 1e-9
        return
    for plan in (fe.mixed_r_plan(arr), None):
        R, vals, xi = fe.assemble_element_block_mixed(mat, nw, arr, Ut, xp, stab_mult=1.0, r_plan=plan, U_prev=Upt)
        torch.cuda.synchronize()
        v, Rn = vals.cpu().numpy(), R.cpu().numpy()
        nu, npd = 3 * n_b, n_b
        sizes = np.cumsum([0, n_e * nu * nu, n_e * nu * npd, n_e * npd * nu, n_e * npd * npd])
        assert rel_err(xi.cpu().numpy(), g["xi"]) < 1e-10, (case, "xi")
        assert rel_err(Rn[eq_u], g["R_u"].reshape(n_e, -1)) < 1e-9, (case, "R_u")
        assert rel_err(Rn[eq_p], g["R_p"]) < 1e-9, (case, "R_p")
        for k, (lo, hi), shp in (("K_uu", sizes[0:2], (n_e, nu, nu)), ("K_up", sizes[1:3], (n_e, nu, npd)),
                                 ("K_pu", sizes[2:4], (n_e, npd, nu)), ("K_pp", sizes[3:5], (n_e, npd, npd))):
            assert rel_err(v[lo:hi].reshape(shp), g[k]) < 1e-9, (case, k, rel_err(v[lo:hi].reshape(shp), g[k]))
    R2, none, _ = fe.assemble_element_block_mixed(mat, nw, arr, Ut, xp, stab_mult=1.0, want_K=False, U_prev=Upt)
    assert none is None and rel_err(R2.cpu().numpy(), Rn) < 1e-12
