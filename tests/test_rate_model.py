"""SmallRateElasticPlastic (cmad/models/small_rate_elastic_plastic.py:34-100, 250-346), the
rate form of the small-strain model (state = [cauchy(6), alpha]), FULL_3D - against golden
vectors produced by executing the reference's own source (make_reference_golden.py section F,
fixture ref_rate_model.npz): the `Model` object under the imperative `newton_solve` along a
two-leg history ((ii, ||C||), xi, Sigma, dC/dxi, dC/dxi_prev, dC/dp) and `make_newton_solve`
+ its IFT rule from the same previous states.

CPU: the torch-AD oracle's restatement (`cmad_oracle.rate_residual`) vs the fixtures.
GPU: `mp_update_rate.cu` through the C-ABI vs the fixtures (values 1e-10, derivatives 1e-9,
iteration counts exactly equal).
"""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import active_all_scalars, active_kernel_set, const_like, material
from tests.helpers import UP, rel_err

from tests.test_def_types import _Fixtures  # noqa: E402

# ref_rate_model_barlat.npz / ref_rate_objectives_barlat.npz: the same jobs with the Yld2004-18p surface
# (`make_reference_golden.py --only barlat_more`)
RT = _Fixtures("ref_rate_model.npz", "ref_rate_model_barlat.npz")
KINDS = sorted({k.split(".")[0] for k in RT.files})


@pytest.mark.parametrize("kind", KINDS)
def test_torch_oracle_vs_reference_rate_model(kind):
    import torch
    from oracle import cmad_oracle as co
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    tv = co.to_torch_tree(material(kind))
    F = RT[f"{kind}.F"]
    x = torch.zeros(7, dtype=torch.float64)
    for t in range(1, 13):
        gu = torch.from_numpy(F[:, :, t] - np.eye(3)); gup = torch.from_numpy(F[:, :, t - 1] - np.eye(3))
        xn, info = co.newton_imperative(x, tv, gu, gup, spec)
        assert info.iters == RT[f"{kind}.iters"][t - 1], (kind, t)
        assert rel_err(xn.numpy(), RT[f"{kind}.xi"][t - 1]) < 1e-10
        assert rel_err(co.dC_dxi(xn, x, tv, gu, gup, spec).numpy(), RT[f"{kind}.dC_dxi"][t - 1]) < 1e-9
        x = xn
    assert RT[f"{kind}.xi"][:, 6].max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_vs_reference_rate_model(cuda_device, kind):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    values = material(kind)
    P = Parameters(values, active_kernel_set(values), const_like(values, None))
    mat = material_from_values(values, model="small_rate_elastic_plastic")
    pid, aidx = active_param_ids(P), np.asarray(P.active_idx)
    F = RT[f"{kind}.F"]
    N = F.shape[2] - 1
    want = ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp", "dxi_deps", "dsig_deps")
    cols = [((3 * k + l,) if k == l else (3 * k + l, 3 * l + k)) for k, l in UP]
    xi = torch.zeros((7, 1), dtype=torch.float64, device=cuda_device)
    for t in range(1, N + 1):
        dgu = (F[:, :, t] - F[:, :, t - 1]).reshape(9, 1)                 # the kernel takes the INCREMENT
        e = torch.from_numpy(dgu.copy()).to(cuda_device)
        ot = mp.mp_update(mat, NewtonSettings(mode="traced"), pid, xi, e, outputs=want)
        assert int(ot["iters"][0]) == RT[f"{kind}.traced_iters"][t - 1], (kind, t, "traced count")
        assert rel_err(ot["xi"][:, 0].cpu().numpy(), RT[f"{kind}.traced_xi"][t - 1]) < 1e-10
        dx = np.stack([sum(RT[f"{kind}.dxi_dgradu"][t - 1][:, c] for c in cc) for cc in cols], axis=-1)   # (7, 6)
        assert rel_err(ot["dxi_deps"][:, 0].cpu().numpy().reshape(7, 6), dx) < 1e-8, (kind, t, "dxi_deps")
        assert rel_err(ot["dsig_deps"][:, 0].cpu().numpy().reshape(6, 6), dx[:6]) < 1e-8
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), pid, xi, e, outputs=want)
        assert int(o["iters"][0]) == RT[f"{kind}.iters"][t - 1], (kind, t)
        assert abs(float(o["cnorm"][0]) - RT[f"{kind}.cnorm"][t - 1]) < 1e-11
        assert rel_err(o["xi"][:, 0].cpu().numpy(), RT[f"{kind}.xi"][t - 1]) < 1e-10, (kind, t)
        assert rel_err(o["sigma"][:, 0].cpu().numpy(), RT[f"{kind}.sigma"][t - 1]) < 1e-10
        assert rel_err(o["dC_dxi"][:, 0].cpu().numpy().reshape(7, 7), RT[f"{kind}.dC_dxi"][t - 1]) < 1e-9
        assert rel_err(o["dC_dxi_prev"][:, 0].cpu().numpy().reshape(7, 7), RT[f"{kind}.dC_dxi_prev"][t - 1]) < 1e-9
        assert rel_err(o["dC_dp"][:, 0].cpu().numpy().reshape(7, len(aidx)),
                       RT[f"{kind}.dC_dp"][t - 1][:, aidx]) < 1e-9, (kind, t, "dC_dp")
        xi = o["xi"]


# ------------------------------------------------------------------------------------------ #
#  Calibration objectives over the rate form: MPAdjointObjective / MPDirectObjective          #
#  (mp_objective.py:92-215) run by the reference on SmallRateElasticPlastic                    #
#  (make_reference_golden.py, `rate_objective`; fixture ref_rate_objectives.npz)               #
# ------------------------------------------------------------------------------------------ #
RO = _Fixtures("ref_rate_objectives.npz", "ref_rate_objectives_barlat.npz")
RO_CASES = sorted({k.rsplit(".", 1)[0] for k in RO.files})


def _rate_objective_parameters(case):
    from tests.golden.materials import objective_trees
    kind, mode = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    return values, act, tr


def test_reference_adjoint_equals_direct_for_the_rate_form():
    for case in RO_CASES:
        assert abs(RO[f"{case}.J_adjoint"] - RO[f"{case}.J_direct"]) < 1e-12 * abs(RO[f"{case}.J_direct"])
        assert rel_err(RO[f"{case}.grad_adjoint"], RO[f"{case}.grad_direct"]) < 1e-9, case


@pytest.mark.parametrize("case", ["J2.native", "hosford.scaled"])
def test_torch_oracle_rate_objective_vs_reference(case):
    from oracle import cmad_oracle as co
    values, act, tr = _rate_objective_parameters(case)
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    J, g = co.mp_objective_adjoint(P, RO[f"{case}.F"], RO[f"{case}.data"], RO[f"{case}.weight"], spec,
                                   RO[f"{case}.x_canonical"], True)
    assert abs(J - RO[f"{case}.J_adjoint"]) < 1e-11 * abs(J)
    assert rel_err(g, RO[f"{case}.grad_adjoint"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", RO_CASES)
def test_cuda_rate_objectives_vs_reference(cuda_device, case):
    """K1-rate per step (the history carries total strains, the increment is formed on the device)
    + K2-rate (mp_sens_rate.cu), through the reference's constructor signatures."""
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectObjective, SmallRateElasticPlastic
    values, act, tr = _rate_objective_parameters(case)
    for strategy, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        P = Parameters(values, act, tr)
        assert np.array_equal(P.active_idx, RO[f"{case}.active_idx"])
        obj = ctor(Calibration(SmallRateElasticPlastic(P), RO[f"{case}.data"], RO[f"{case}.weight"]),
                   RO[f"{case}.F"], device=cuda_device)
        r = obj.evaluate(RO[f"{case}.x_canonical"])
        assert abs(r.J - RO[f"{case}.J_{strategy}"]) < 1e-11 * abs(r.J), (case, strategy)
        assert rel_err(r.grad, RO[f"{case}.grad_{strategy}"]) < 1e-9, (case, strategy, r.grad)


# ------------------------------------------------------------------------------------------ #
#  FE element blocks of the rate form: per_element_R_and_K_coupled over the rate model's        #
#  per-IP COUPLED evaluator, with the real previous displacement (fixture ref_rate_fe_elements) #
# ------------------------------------------------------------------------------------------ #
RF = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rate_fe_elements.npz"))
RF_CASES = sorted({k.rsplit(".", 1)[0] for k in RF.files})
FE_NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)          # for_model's COUPLED defaults


def _rate_fe_arrays(case):
    g = {k: RF[f"{case}.{k}"] for k in ("U", "U_prev", "xi_prev", "grad_N", "det", "xi", "R_u", "K_uu", "R_only_u")}
    n_e, n_b = g["U"].shape[0], g["U"].shape[1]
    conn = np.arange(n_e * n_b).reshape(n_e, n_b)
    eq_u = (conn[:, :, None] * 3 + np.arange(3)[None, None, :]).reshape(n_e, 3 * n_b)
    return g, eq_u, RF[f"{case}.quad_w"], RF[f"{case}.N"]


def test_rate_fe_fixture_is_plastic_and_history_dependent():
    for case in RF_CASES:
        g, *_ = _rate_fe_arrays(case)
        assert g["xi"][..., 6].max() > 0 and np.abs(g["U_prev"]).max() > 0
        assert rel_err(g["R_only_u"], g["R_u"]) < 1e-12


@pytest.mark.parametrize("case", ["tet4.J2", "hex8.hosford"])
def test_torch_oracle_rate_elements_vs_reference(case):
    import torch
    from oracle import cmad_oracle as co
    family, kind = case.split(".")
    g, eq_u, quad_w, N = _rate_fe_arrays(case)
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    tv = co.to_torch_tree(material(kind))
    for e in range(0, g["U"].shape[0], 3):                      # every third record: torch AD per element is slow
        R, K, xi, _ = co.coupled_element(tv, torch.from_numpy(g["U"][e]), torch.from_numpy(g["U_prev"][e]),
                                      torch.from_numpy(g["xi_prev"][e]), torch.from_numpy(g["grad_N"][e]),
                                      torch.from_numpy(g["det"][e]), torch.from_numpy(quad_w), spec,
                                      newton_settings=FE_NEWTON)
        n_b = g["U"].shape[1]
        assert rel_err(np.asarray(xi), g["xi"][e]) < 1e-10, (case, e)
        assert rel_err(np.asarray(R).reshape(-1), g["R_u"][e].reshape(-1)) < 1e-9
        assert rel_err(np.asarray(K).reshape(3 * n_b, 3 * n_b), g["K_uu"][e]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", RF_CASES)
def test_cuda_rate_fe_blocks_vs_reference(cuda_device, case):
    import torch
    from cmad_b200 import fe, material_from_values
    from cmad_b200.fe_mesh import FEBlockArrays
    family, kind = case.split(".")
    g, eq_u, quad_w, N = _rate_fe_arrays(case)
    n_e, n_b = eq_u.shape[0], eq_u.shape[1] // 3
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(cuda_device)  # noqa: E731
    arr = FEBlockArrays(t(eq_u, torch.int32), t(g["grad_N"]), t(g["det"]), t(quad_w), t(N), int(g["U"].size), None, None)
    mat = material_from_values(material(kind), model="small_rate_elastic_plastic")
    nw = fe.fe_newton_settings(**FE_NEWTON)
    U, Up, xp = t(g["U"].reshape(-1)), t(g["U_prev"].reshape(-1)), t(g["xi_prev"])
    plan = fe.SegmentPlan(eq_u.reshape(-1), int(g["U"].size), device=cuda_device)
    R, vals, xi = fe.assemble_element_block(mat, nw, arr, U, xp, r_plan=plan, U_prev=Up)
    torch.cuda.synchronize()
    assert rel_err(xi.cpu().numpy(), g["xi"]) < 1e-10, (case, "xi")
    assert rel_err(R.cpu().numpy()[eq_u], g["R_u"].reshape(n_e, -1)) < 1e-9, (case, "R")
    assert rel_err(vals.cpu().numpy().reshape(n_e, 3 * n_b, 3 * n_b), g["K_uu"]) < 1e-9, (case, "K")
    # residual-only launch (K4) and the missing-U_prev error
    o = fe.fe_block_launch(mat, nw, arr, U, xp, ("xi", "R_elem"), U_prev=Up)
    assert rel_err(o["R_elem"].cpu().numpy(), g["R_u"].reshape(n_e, -1)) < 1e-9
    with pytest.raises(ValueError):
        fe.fe_block_launch(mat, nw, arr, U, xp, ("xi", "R_elem"))
