"""SmallRateElasticPlastic under the PLANE_STRESS / UNIAXIAL_STRESS deformation types
(cmad/models/small_rate_elastic_plastic.py:34-77, 296-345; the "small rate" cases of the reference's
tests/models/test_elastic_plastic_models.py): n_xi = 8 (out-of-plane stretch) / 12 (two stretches +
three off-axis delta strains), constraints on the GLOBAL stress increment.  Fixture
tests/golden/ref_def_types_rate.npz: the reference's own `Model` object under the imperative
`newton_solve` and `make_newton_solve` + its IFT rule, with the AD products dC/dxi, dC/dxi_prev,
dC/dp - J2, rotated anisotropic Hill, Hosford (`make_reference_golden.py --only rate_deftypes`).

CPU: the torch-AD oracle.  GPU: mp_update_rate_dt.cu through the C-ABI - counts exact, states 1e-10,
derivatives 1e-9."""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import active_kernel_set, const_like, material
from tests.helpers import UP, rel_err
from tests.test_def_types import DEF, _Fixtures, _strain_rows, _sym_cols_2d

# ref_def_types_rate_barlat.npz: the same jobs with the Yld2004-18p surface (`--only rate_deftypes_barlat`)
RD = _Fixtures("ref_def_types_rate.npz", "ref_def_types_rate_barlat.npz")
CASES = sorted({".".join(k.split(".")[:2]) for k in RD.files})


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_vs_reference_rate_def_types(case):
    import torch
    from oracle import cmad_oracle as co
    kind, dtn = case.split(".")
    spec = co.ModelSpec(kind="small_rate_elastic_plastic", def_type=getattr(co, dtn))
    assert spec.num_dofs == (8 if dtn == "PLANE_STRESS" else 12)
    tv = co.to_torch_tree(material(kind))
    F = RD[f"{case}.F"]
    nd = F.shape[0]
    x = torch.as_tensor(spec.init_xi())
    for t in range(1, 8):
        gu, gup = torch.from_numpy(F[:, :, t] - np.eye(nd)), torch.from_numpy(F[:, :, t - 1] - np.eye(nd))
        xn, info = co.newton_imperative(x, tv, gu, gup, spec)
        assert info.iters == RD[f"{case}.iters"][t - 1], (case, t)
        assert rel_err(xn.numpy(), RD[f"{case}.xi"][t - 1]) < 1e-10
        assert rel_err(co.dC_dxi(xn, x, tv, gu, gup, spec).numpy(), RD[f"{case}.dC_dxi"][t - 1]) < 1e-9
        assert rel_err(co.dC_dxi_prev(xn, x, tv, gu, gup, spec).numpy(), RD[f"{case}.dC_dxi_prev"][t - 1]) < 1e-9
        x = xn
    assert RD[f"{case}.xi"][:, 6].max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_vs_reference_rate_def_types(cuda_device, case):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    kind, dtn = case.split(".")
    dt = DEF[dtn]
    values = material(kind)
    P = Parameters(values, active_kernel_set(values), const_like(values, None))
    mat, pid = material_from_values(values, model="small_rate_elastic_plastic"), active_param_ids(P)
    aidx = np.asarray(P.active_idx)
    F = RD[f"{case}.F"]
    nd, N = F.shape[0], F.shape[2] - 1
    rows = _strain_rows(case, F)
    nxi, ns = (8, 3) if dt == 1 else (12, 1)
    assert mp.n_xi_of(mat, dt) == nxi
    want = ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp", "dxi_deps", "dsig_deps", "flags")
    xi = mp.init_xi(mat, 1, cuda_device, def_type=dt)
    row9 = [3 * i + j for i, j in UP]
    for t in range(1, N + 1):
        e = torch.from_numpy((rows[:, t:t + 1] - rows[:, t - 1:t]).copy()).to(cuda_device)       # the INCREMENT
        ot = mp.mp_update(mat, NewtonSettings(mode="traced"), pid, xi, e, outputs=want, def_type=dt)
        assert int(ot["iters"][0]) == RD[f"{case}.traced_iters"][t - 1], (case, t, "traced count")
        assert rel_err(ot["xi"][:, 0].cpu().numpy(), RD[f"{case}.traced_xi"][t - 1]) < 1e-10
        dx = _sym_cols_2d(RD[f"{case}.dxi_dgradu"][t - 1], nd)
        ds = _sym_cols_2d(RD[f"{case}.dsig_dgradu"][t - 1], nd)[row9]
        assert rel_err(ot["dxi_deps"][:, 0].cpu().numpy().reshape(nxi, ns), dx) < 1e-8, (case, t, "dxi_deps")
        assert rel_err(ot["dsig_deps"][:, 0].cpu().numpy().reshape(6, ns), ds) < 1e-8, (case, t, "dsig_deps")
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), pid, xi, e, outputs=want, def_type=dt)
        assert int(o["iters"][0]) == RD[f"{case}.iters"][t - 1], (case, t)
        assert abs(float(o["cnorm"][0]) - RD[f"{case}.cnorm"][t - 1]) < 1e-11
        assert rel_err(o["xi"][:, 0].cpu().numpy(), RD[f"{case}.xi"][t - 1]) < 1e-10, (case, t)
        assert rel_err(o["sigma"][:, 0].cpu().numpy(), RD[f"{case}.sigma"][t - 1]) < 1e-10, (case, t)
        assert rel_err(o["dC_dxi"][:, 0].cpu().numpy().reshape(nxi, nxi), RD[f"{case}.dC_dxi"][t - 1]) < 1e-9
        assert rel_err(o["dC_dxi_prev"][:, 0].cpu().numpy().reshape(nxi, nxi), RD[f"{case}.dC_dxi_prev"][t - 1]) < 1e-9
        assert rel_err(o["dC_dp"][:, 0].cpu().numpy().reshape(nxi, len(aidx)),
                       RD[f"{case}.dC_dp"][t - 1][:, aidx]) < 1e-9, (case, t, "dC_dp")
        xi = o["xi"]
    s = o["sigma"][:, 0].cpu().numpy()
    assert abs(s[5]) < 1e-7 and (dt == 1 or abs(s[3]) < 1e-7)


@pytest.mark.gpu
def test_cuda_rate_def_type_batch(cuda_device):
    """A batch of points on scaled copies of the fixture's path: the plane-stress constraint holds for
    every point and every point yields."""
    import torch
    from cmad_b200 import NewtonSettings, material_from_values, mp
    case, dt = "hill_rot.PLANE_STRESS", 1
    mat = material_from_values(material("hill_rot"), model="small_rate_elastic_plastic")
    rows = _strain_rows(case, RD[f"{case}.F"])
    n = 64
    scale = torch.linspace(0.5, 1.5, n, dtype=torch.float64, device=cuda_device)
    xi = mp.init_xi(mat, n, cuda_device, def_type=dt)
    for t in range(1, 12):
        e = torch.from_numpy((rows[:, t] - rows[:, t - 1]).copy()).to(cuda_device)[:, None] * scale[None, :]
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), np.zeros(0, np.int32), xi, e.contiguous(),
                         outputs=("xi", "sigma", "iters"), def_type=dt)
        xi = o["xi"]
    assert float(o["sigma"][5].abs().max()) < 1e-7
    assert float(xi[6].min()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["J2.UNIAXIAL_STRESS", "hill_rot.PLANE_STRESS"])
def test_cuda_primal_loop_rate_def_types(cuda_device, case):
    """`cmad primal` on the rate form under a def-type: the Python stand-in + run_primal_pass reproduce
    the reference's trajectory (block layout [6, 1, 1] / [6, 1, 2, 3])."""
    from cmad_b200 import objectives as ob
    from cmad_b200.primal import block_sizes, run_primal_pass
    kind, dtn = case.split(".")
    values = material(kind)
    P = Parameters(values, const_like(values, False), const_like(values, None))
    model = ob.SmallRateElasticPlastic(P, def_type=getattr(ob, dtn))
    assert block_sizes(model) == ([6, 1, 1] if dtn == "PLANE_STRESS" else [6, 1, 2, 3])
    F = RD[f"{case}.F"]
    N = F.shape[2] - 1
    cauchy, xi_traj, log, _ = run_primal_pass(model, F, N, device=cuda_device)
    assert [s["iters"] for s in log] == list(RD[f"{case}.iters"])
    xi_end = np.concatenate([np.ravel(b) for b in xi_traj[-1]])
    assert rel_err(xi_end, RD[f"{case}.xi"][-1]) < 1e-9
    sig = np.array([cauchy[i, j, -1] for i, j in UP])
    assert rel_err(sig, RD[f"{case}.sigma"][-1]) < 1e-9


def test_reference_adjoint_equals_direct():
    for case in CASES:
        for tag in ("scaled", "native"):
            a, d = RD[f"{case}.obj_{tag}.grad_adjoint"], RD[f"{case}.obj_{tag}.grad_direct"]
            assert rel_err(a, d) < 1e-9, (case, tag)


@pytest.mark.parametrize("case", ["J2.PLANE_STRESS", "hill_rot.UNIAXIAL_STRESS"])
def test_torch_oracle_rate_def_type_objective_vs_reference(case):
    from oracle import cmad_oracle as co
    from tests.golden.materials import objective_trees
    kind, dtn = case.split(".")
    pre = f"{case}.obj_native"
    values, act, tr = objective_trees(kind, False)
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec(kind="small_rate_elastic_plastic", def_type=getattr(co, dtn))
    J, g = co.mp_objective_adjoint(P, RD[f"{case}.F"], RD[f"{pre}.data"], RD[f"{pre}.weight"], spec,
                                   RD[f"{pre}.x_canonical"], True)
    assert abs(J - RD[f"{pre}.J_adjoint"]) < 1e-10 * abs(J)
    assert rel_err(np.asarray(g), RD[f"{pre}.grad_adjoint"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["scaled", "native"])
def test_cuda_objectives_vs_reference_rate_def_types(cuda_device, case, tag):
    """The KA5 setting on the rate form (tests/objectives/test_J2_fd_checks.py runs its plane-stress
    gradient checks on both small-strain models): forward history (K1 per step, increments formed on
    the device) + K2 adjoint / direct on the bordered systems, reference constructor signatures."""
    from cmad_b200 import objectives as ob
    from tests.golden.materials import objective_trees
    kind, dtn = case.split(".")
    F = RD[f"{case}.F"]
    pre = f"{case}.obj_{tag}"
    for strategy, ctor in (("adjoint", ob.MPAdjointObjective), ("direct", ob.MPDirectObjective)):
        values, act, tr = objective_trees(kind, tag == "scaled")
        P = Parameters(values, act, tr)
        assert np.array_equal(P.active_idx, RD[f"{pre}.active_idx"])
        model = ob.SmallRateElasticPlastic(P, def_type=getattr(ob, dtn))
        obj = ctor(ob.Calibration(model, RD[f"{pre}.data"], RD[f"{pre}.weight"]), F, device=cuda_device)
        r = obj.evaluate(RD[f"{pre}.x_canonical"])
        assert abs(r.J - RD[f"{pre}.J_{strategy}"]) < 1e-10 * abs(r.J), (case, tag, strategy)
        assert rel_err(r.grad, RD[f"{pre}.grad_{strategy}"]) < 1e-8, (case, tag, strategy, r.grad)


# ------------------------------------------------------------------------------------------ #
#  UniaxialCalibration on the rate form (axial stress + the two off-axis stretches, per-step    #
#  weights; stretch block 2 of the rate model as well): ref_uniaxial_qoi_rate.npz               #
# ------------------------------------------------------------------------------------------ #
_UQR = os.path.join(os.path.dirname(__file__), "golden", "ref_uniaxial_qoi_rate.npz")
UQR = np.load(_UQR)
UQR_CASES = sorted({k.rsplit(".", 1)[0] for k in UQR.files})


def test_uniaxial_qoi_rate_fixture_is_consistent():
    assert {c.split(".")[0] for c in UQR_CASES} == {"J2", "hill_rot", "hosford"}
    for case in UQR_CASES:
        assert abs(UQR[f"{case}.J_adjoint"] - UQR[f"{case}.J_direct"]) < 1e-12 * abs(UQR[f"{case}.J_direct"])
        assert rel_err(UQR[f"{case}.grad_adjoint"], UQR[f"{case}.grad_direct"]) < 1e-8
        assert np.abs(UQR[f"{case}.data"][1:]).max() > 1e-3           # lateral strains are measured


@pytest.mark.parametrize("case", ["J2.native", "hill_rot.scaled"])
def test_torch_oracle_uniaxial_qoi_rate_vs_reference(case):
    from oracle import cmad_oracle as co
    from tests.golden.materials import objective_trees
    kind, tag = case.split(".")
    P = co.OracleParameters(*objective_trees(kind, tag == "scaled"))
    spec = co.ModelSpec(kind="small_rate_elastic_plastic", def_type=co.UNIAXIAL_STRESS)
    J, g = co.mp_objective_adjoint(P, UQR[f"{case}.F"], UQR[f"{case}.data"], UQR[f"{case}.weight"], spec,
                                   UQR[f"{case}.x_canonical"], True)
    assert abs(J - UQR[f"{case}.J_adjoint"]) < 1e-10 * abs(J)
    assert rel_err(np.asarray(g), UQR[f"{case}.grad_adjoint"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", UQR_CASES)
def test_cuda_uniaxial_qoi_rate_vs_reference(cuda_device, case):
    from cmad_b200 import objectives as ob
    from tests.golden.materials import objective_trees
    kind, tag = case.split(".")
    for strategy, ctor in (("adjoint", ob.MPAdjointObjective), ("direct", ob.MPDirectObjective)):
        P = Parameters(*objective_trees(kind, tag == "scaled"))
        assert np.array_equal(P.active_idx, UQR[f"{case}.active_idx"])
        model = ob.SmallRateElasticPlastic(P, def_type=ob.UNIAXIAL_STRESS)
        qoi = ob.UniaxialCalibration(model, UQR[f"{case}.data"], UQR[f"{case}.weight"], uniaxial_stress_idx=0, stretch_var_idx=2)
        r = ctor(qoi, UQR[f"{case}.F"], device=cuda_device).evaluate(UQR[f"{case}.x_canonical"])
        assert abs(r.J - UQR[f"{case}.J_{strategy}"]) < 1e-10 * abs(r.J), (case, strategy)
        assert rel_err(r.grad, UQR[f"{case}.grad_{strategy}"]) < 1e-8, (case, strategy, r.grad)
