"""PLANE_STRESS / UNIAXIAL_STRESS deformation types of SmallElasticPlastic
(small_elastic_plastic.py:126-180, 274-302): n_xi = 8 / 9 with the stretch unknowns and the
stress-constraint rows, against golden vectors produced by executing the reference's own
source (tests/golden/make_reference_golden.py section E, fixture ref_def_types.npz):
the `Model` object under the imperative `newton_solve` along two-leg histories (xi, Sigma,
(ii, ||C||), dC/dxi, dC/dxi_prev, dC/dp), `make_newton_solve` + its IFT rule from the same
previous states (xi, count, d(xi, sigma)/d(grad_u)), and the KA5-style adjoint / direct
objective gradients (tests/objectives/test_J2_fd_checks.py:303-349).

CPU: the torch-AD oracle (the only oracle that carries the def-types) vs the fixtures.
GPU: the CUDA kernels (mp_update_dt.cu; K2 with the bordered systems) vs the fixtures.
"""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import active_all_scalars, active_kernel_set, const_like, material, objective_trees
from tests.helpers import UP, rel_err

class _Fixtures:
    """Several .npz fixtures behind one lookup (keys are disjoint: `<kind>.<def type>.<array>`)."""

    def __init__(self, *names):
        paths = [os.path.join(os.path.dirname(__file__), "golden", n) for n in names]
        assert os.path.exists(paths[0]), paths[0]
        # later files extend the first with more cases; one that has not been generated adds none
        self._z = [np.load(p) for p in paths if os.path.exists(p)]
        self.files = [k for z in self._z for k in z.files]

    def __getitem__(self, key):
        for z in self._z:
            if key in z.files:
                return z[key]
        raise KeyError(key)


# ref_def_types_rot.npz: the same jobs on the rotated anisotropic Hill material (`--only deftypes_rot`):
# constraints in global axes, update in material axes (small_elastic_plastic.py:44-62, 287-302)
# ref_def_types_barlat.npz: the Yld2004-18p surface in both def-types (`--only barlat_more`)
DT = _Fixtures("ref_def_types.npz", "ref_def_types_rot.npz", "ref_def_types_barlat.npz")
CASES = sorted({".".join(k.split(".")[:2]) for k in DT.files})
DEF = {"PLANE_STRESS": 1, "UNIAXIAL_STRESS": 2}            # CMADX_DEF_*


def _strain_rows(case, F):
    """prescribed kinematics as the kernel's `strain` rows: 2x2 grad_u row-major / axial strain"""
    nd = F.shape[0]
    gu = F - np.eye(nd)[:, :, None]
    return gu.reshape(nd * nd, -1)                          # (4 | 1, N+1)


def _sym_cols_2d(M, nd):
    """(..., nd*nd) derivative w.r.t. grad_u entries -> (..., ns) w.r.t. symmetric components"""
    if nd == 1:
        return M
    return np.stack([M[..., 0], M[..., 1] + M[..., 2], M[..., 3]], axis=-1)


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_vs_reference_def_types(case):
    import torch
    from oracle import cmad_oracle as co
    kind, dtn = case.split(".")
    values = material(kind)
    spec = co.ModelSpec(def_type=getattr(co, dtn))
    tv = co.to_torch_tree(values)
    F = DT[f"{case}.F"]
    nd = F.shape[0]
    x = torch.as_tensor(spec.init_xi())
    for t in range(1, 9):                                   # first steps: elastic, then plastic
        gu = torch.from_numpy(F[:, :, t] - np.eye(nd))
        gup = torch.from_numpy(F[:, :, t - 1] - np.eye(nd))
        xn, info = co.newton_imperative(x, tv, gu, gup, spec)
        assert info.iters == DT[f"{case}.iters"][t - 1]
        assert rel_err(xn.numpy(), DT[f"{case}.xi"][t - 1]) < 1e-10
        A = co.dC_dxi(xn, x, tv, gu, gup, spec).numpy()
        assert rel_err(A, DT[f"{case}.dC_dxi"][t - 1]) < 1e-9
        x = xn
    assert DT[f"{case}.xi"][:, 6].max() > 0


# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_vs_reference_def_types(cuda_device, case):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    kind, dtn = case.split(".")
    dt = DEF[dtn]
    values = material(kind)
    P = Parameters(values, active_kernel_set(values), const_like(values, None))
    mat, pid = material_from_values(values), active_param_ids(P)
    aidx = np.asarray(P.active_idx)
    F = DT[f"{case}.F"]
    nd, N = F.shape[0], F.shape[2] - 1
    rows = _strain_rows(case, F)
    nxi = 8 if dt == 1 else 9
    ns = 3 if dt == 1 else 1
    want = ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp", "dxi_deps", "dsig_deps", "flags")
    xi = mp.init_xi(mat, 1, cuda_device, def_type=dt)
    row9 = [3 * i + j for i, j in UP]
    for t in range(1, N + 1):
        e = torch.from_numpy(rows[:, t:t + 1].copy()).to(cuda_device)
        # traced flavour (make_newton_solve + IFT rule) from the same previous state
        ot = mp.mp_update(mat, NewtonSettings(mode="traced"), pid, xi, e, outputs=want, def_type=dt)
        assert int(ot["iters"][0]) == DT[f"{case}.traced_iters"][t - 1], (case, t, "traced count")
        assert rel_err(ot["xi"][:, 0].cpu().numpy(), DT[f"{case}.traced_xi"][t - 1]) < 1e-10
        dx = _sym_cols_2d(DT[f"{case}.dxi_dgradu"][t - 1], nd)                      # (nxi, ns)
        ds = _sym_cols_2d(DT[f"{case}.dsig_dgradu"][t - 1], nd)[row9]               # (6, ns)
        assert rel_err(ot["dxi_deps"][:, 0].cpu().numpy().reshape(nxi, ns), dx) < 1e-8, (case, t, "dxi_deps")
        assert rel_err(ot["dsig_deps"][:, 0].cpu().numpy().reshape(6, ns), ds) < 1e-8, (case, t, "dsig_deps")
        # imperative flavour (the Model object under newton_solve)
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), pid, xi, e, outputs=want, def_type=dt)
        assert int(o["iters"][0]) == DT[f"{case}.iters"][t - 1], (case, t)
        assert abs(float(o["cnorm"][0]) - DT[f"{case}.cnorm"][t - 1]) < 1e-11
        assert rel_err(o["xi"][:, 0].cpu().numpy(), DT[f"{case}.xi"][t - 1]) < 1e-10, (case, t)
        assert rel_err(o["sigma"][:, 0].cpu().numpy(), DT[f"{case}.sigma"][t - 1]) < 1e-10, (case, t)
        assert rel_err(o["dC_dxi"][:, 0].cpu().numpy().reshape(nxi, nxi), DT[f"{case}.dC_dxi"][t - 1]) < 1e-9
        assert rel_err(o["dC_dxi_prev"][:, 0].cpu().numpy().reshape(nxi, nxi), DT[f"{case}.dC_dxi_prev"][t - 1]) < 1e-9
        assert rel_err(o["dC_dp"][:, 0].cpu().numpy().reshape(nxi, len(aidx)),
                       DT[f"{case}.dC_dp"][t - 1][:, aidx]) < 1e-9, (case, t, "dC_dp")
        xi = o["xi"]
    # stress constraint of the def-type at the end state
    s = o["sigma"][:, 0].cpu().numpy()
    assert abs(s[5]) < 1e-8 and (dt == 1 or abs(s[3]) < 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("dtn", ["PLANE_STRESS", "UNIAXIAL_STRESS"])
def test_cuda_dC_dp_hosford_exponent_def_types(cuda_device, dtn):
    """The `hosford a` column of dC/dp (the reference's jacrev over ALL leaves, model.py:126-133) from
    the def-type kernel, at the reference's own states (Model.evaluate semantics: max_iters = 0)."""
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    case, dt = f"hosford.{dtn}", DEF[dtn]
    values = material("hosford")
    act = const_like(values, False)
    act["plastic"]["effective stress"]["hosford"]["a"] = True
    P = Parameters(values, act, const_like(values, None))
    aidx = np.asarray(P.active_idx)
    F = DT[f"{case}.F"]
    rows, nxi = _strain_rows(case, F), (8 if dt == 1 else 9)
    mat = material_from_values(values)
    xi_hist = np.vstack([mp.init_xi(mat, 1, "cpu", def_type=dt).numpy().T, DT[f"{case}.xi"]])     # (N+1, nxi)
    n_plastic = 0
    for t in range(1, F.shape[2]):
        xp = torch.from_numpy(xi_hist[t - 1:t].T.copy()).to(cuda_device)
        x = torch.from_numpy(xi_hist[t:t + 1].T.copy()).to(cuda_device)
        e = torch.from_numpy(rows[:, t:t + 1].copy()).to(cuda_device)
        o = mp.mp_update(mat, NewtonSettings(mode="imperative", max_iters=0), active_param_ids(P), xp, e,
                         outputs=("xi", "dC_dp", "flags"), xi_init=x, def_type=dt)
        ref = DT[f"{case}.dC_dp"][t - 1][:, aidx]
        got = o["dC_dp"][:, 0].cpu().numpy().reshape(nxi, 1)
        # uniaxial stress: phi = |sigma_axial| for every exponent, the column is zero up to rounding
        # (1e-35 in the reference's AD, 1e-20 in the closed form): absolute floor at the columns' scale
        assert np.abs(got - ref).max() < 1e-9 * max(np.abs(ref).max(), 1e-6), (case, t)
        n_plastic += int(np.abs(ref).max() > 1e-12)
    assert n_plastic > 5 or dtn == "UNIAXIAL_STRESS"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["scaled", "native"])
def test_cuda_objectives_vs_reference_def_types(cuda_device, case, tag):
    """KA5 on the CUDA path: `MPAdjointObjective` / `MPDirectObjective(Calibration(model, data,
    weight), F).evaluate(x)` with the reference's constructor signatures, plane stress and
    uniaxial stress, canonical (log / bounds) and native parameters."""
    from cmad_b200 import objectives as ob
    kind, dtn = case.split(".")
    F = DT[f"{case}.F"]
    pre = f"{case}.obj_{tag}"
    for strategy, ctor in (("adjoint", ob.MPAdjointObjective), ("direct", ob.MPDirectObjective)):
        values, act, tr = objective_trees(kind, tag == "scaled")
        P = Parameters(values, act, tr)
        assert np.array_equal(P.active_idx, DT[f"{pre}.active_idx"])
        model = ob.SmallElasticPlastic(P, def_type=getattr(ob, dtn))
        obj = ctor(ob.Calibration(model, DT[f"{pre}.data"], DT[f"{pre}.weight"]), F, device=cuda_device)
        r = obj.evaluate(DT[f"{pre}.x_canonical"])
        assert abs(r.J - DT[f"{pre}.J_{strategy}"]) < 1e-10 * abs(r.J), (case, tag, strategy)
        assert rel_err(r.grad, DT[f"{pre}.grad_{strategy}"]) < 1e-8, (case, tag, strategy, r.grad)


# ------------------------------------------------------------------------------------------ #
#  UniaxialCalibration (cmad/qois/uniaxial_calibration.py): axial stress + the two off-axis   #
#  strains with per-step weights, run by the reference (make_reference_golden.py section G,   #
#  fixture ref_uniaxial_qoi.npz)                                                               #
# ------------------------------------------------------------------------------------------ #
UQ = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_uniaxial_qoi.npz"))
UQ_CASES = sorted({k.rsplit(".", 1)[0] for k in UQ.files})


def test_uniaxial_qoi_fixture_is_consistent():
    for case in UQ_CASES:
        assert abs(UQ[f"{case}.J_adjoint"] - UQ[f"{case}.J_direct"]) < 1e-12 * abs(UQ[f"{case}.J_direct"])
        assert rel_err(UQ[f"{case}.grad_adjoint"], UQ[f"{case}.grad_direct"]) < 1e-8
        assert UQ[f"{case}.weight"].shape == UQ[f"{case}.data"].shape == (3, UQ[f"{case}.F"].shape[2])
        assert np.abs(UQ[f"{case}.data"][1:]).max() > 1e-3           # lateral strains are measured


@pytest.mark.parametrize("case", ["J2.native", "hosford.scaled"])
def test_torch_oracle_uniaxial_qoi_vs_reference(case):
    from oracle import cmad_oracle as co
    kind, tag = case.split(".")
    values, act, tr = objective_trees(kind, tag == "scaled")
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec(kind="small_elastic_plastic", def_type=co.UNIAXIAL_STRESS)
    J, g = co.mp_objective_adjoint(P, UQ[f"{case}.F"], UQ[f"{case}.data"], UQ[f"{case}.weight"], spec,
                                   UQ[f"{case}.x_canonical"], True)
    assert abs(J - UQ[f"{case}.J_adjoint"]) < 1e-10 * abs(J)
    assert rel_err(g, UQ[f"{case}.grad_adjoint"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", UQ_CASES)
def test_cuda_uniaxial_qoi_vs_reference(cuda_device, case):
    """`MPAdjointObjective / MPDirectObjective(UniaxialCalibration(model, data, weight, 0, 2), F)` on
    the CUDA path (K1-DT + K2-DT with CMADX_QOI_UNIAXIAL_CALIBRATION)."""
    from cmad_b200 import objectives as ob
    kind, tag = case.split(".")
    for strategy, ctor in (("adjoint", ob.MPAdjointObjective), ("direct", ob.MPDirectObjective)):
        values, act, tr = objective_trees(kind, tag == "scaled")
        P = Parameters(values, act, tr)
        assert np.array_equal(P.active_idx, UQ[f"{case}.active_idx"])
        model = ob.SmallElasticPlastic(P, def_type=ob.UNIAXIAL_STRESS)
        qoi = ob.UniaxialCalibration(model, UQ[f"{case}.data"], UQ[f"{case}.weight"], uniaxial_stress_idx=0, stretch_var_idx=2)
        r = ctor(qoi, UQ[f"{case}.F"], device=cuda_device).evaluate(UQ[f"{case}.x_canonical"])
        assert abs(r.J - UQ[f"{case}.J_{strategy}"]) < 1e-10 * abs(r.J), (case, strategy)
        assert rel_err(r.grad, UQ[f"{case}.grad_{strategy}"]) < 1e-8, (case, strategy, r.grad)


@pytest.mark.gpu
def test_cuda_uniaxial_qoi_refusals(cuda_device):
    from cmad_b200 import objectives as ob
    case = "J2.scaled"
    values, act, tr = objective_trees("J2", True)
    P = Parameters(values, act, tr)
    with pytest.raises(NotImplementedError):          # a FULL_3D model has no off-axis stretches
        ob.UniaxialCalibration(ob.SmallElasticPlastic(P), UQ[f"{case}.data"], UQ[f"{case}.weight"])
    model = ob.SmallElasticPlastic(P, def_type=ob.UNIAXIAL_STRESS)
    qoi = ob.UniaxialCalibration(model, UQ[f"{case}.data"], UQ[f"{case}.weight"])
    with pytest.raises(NotImplementedError):          # no second-order pass for this QoI
        ob.MPDirectAdjointObjective(qoi, UQ[f"{case}.F"], device=cuda_device).evaluate(UQ[f"{case}.x_canonical"])
