"""Second-order (Hessian) path of the material-point calibration objective.

Fixture ``tests/golden/ref_mp_hessian.npz``: the reference's own, unmodified
``MPDirectAdjointObjective`` (cmad/objectives/mp_objective.py:218-343, through
``Model.evaluate_hessians`` / ``QoI.evaluate_hessians``) executed by
``tests/golden/make_reference_golden.py --only hessian`` - J, the canonical-coordinate gradient
and Hessian for J2 / Hill / Hosford(a = 4), with log / bounds transforms ("scaled", flow-stress
parameters active) and without ("native", elastic + flow-stress parameters active).

CPU: the torch oracle's restatement of the direct-adjoint recurrence against the fixture.
GPU: K2-H (cmadx_mp_objective_hessian, hyper-dual forward mode) through the reference-style
constructor against the fixture, against the torch oracle on a small batch, and against central
differences of the CUDA adjoint gradient (a size-independent property) on a large one.
Tolerance: the reference's own cross-strategy Hessian tolerance is 1e-8
(tests/objectives/test_jvp_vs_original.py:95-97); entries are compared relative to
sqrt(|H_ii H_jj|) (the parameters' scales differ by six decades in the native cases)."""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from oracle import cmad_oracle as co
from tests.golden.materials import objective_trees

G = os.path.join(os.path.dirname(__file__), "golden")
HS = np.load(os.path.join(G, "ref_mp_hessian.npz"))
CASES = sorted({k.rsplit(".", 1)[0] for k in HS.files})


def hess_err(H, Href):
    d = np.sqrt(np.abs(np.diag(Href)))
    return float((np.abs(H - Href) / np.outer(d, d)).max())


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_hessian_vs_reference(case):
    kind, mode = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    P = co.OracleParameters(values, act, tr)
    assert np.array_equal(P.active_idx, HS[f"{case}.active_idx"])
    spec = co.ModelSpec()
    J, g, H = co.mp_objective_direct_adjoint(P, HS[f"{case}.F"], HS[f"{case}.data"], HS[f"{case}.weight"], spec,
                                             HS[f"{case}.x_canonical"], True, reference_qoi_cross_terms=True)
    assert np.allclose(P.flat_values()[P.active_idx], HS[f"{case}.active_native"], rtol=1e-14)
    assert abs(J - HS[f"{case}.J"]) < 1e-11 * abs(J)
    assert np.abs(g - HS[f"{case}.grad"]).max() < 1e-9 * np.abs(g).max()
    assert hess_err(H, HS[f"{case}.hessian"]) < 1e-8
    assert np.array_equal(H, H.T)


@pytest.mark.parametrize("kind", ["J2", "hosford"])
def test_complete_hessian_is_the_derivative_of_the_gradient(kind):
    """With elastic parameters active the reference's Hessian drops the d2J/dxi dp terms
    (qoi.py:53-55 differentiates w.r.t. xi_prev).  The oracle's default keeps them: it equals
    central differences of the adjoint gradient, the reference-compatible variant does not."""
    case = f"{kind}.native"
    F, data, w, x = (HS[f"{case}.{k}"] for k in ("F", "data", "weight", "x_canonical"))
    F, data = F[:, :, :7], data[:, :, :7]
    spec = co.ModelSpec()
    mk = lambda: co.OracleParameters(*objective_trees(kind, False))
    _, _, H = co.mp_objective_direct_adjoint(mk(), F, data, w, spec, x, True)
    _, _, Href = co.mp_objective_direct_adjoint(mk(), F, data, w, spec, x, True, reference_qoi_cross_terms=True)
    fd = np.zeros_like(H)
    for c in range(len(x)):
        h = 1e-6 * abs(x[c])
        gs = []
        for sgn in (1.0, -1.0):
            xx = x.copy(); xx[c] += sgn * h
            gs.append(co.mp_objective_adjoint(mk(), F, data, w, spec, xx, True)[1])
        fd[:, c] = (gs[0] - gs[1]) / (2 * h)
    assert hess_err(H, fd) < 1e-6
    assert hess_err(Href, fd) > 1e-2
    # the flow-stress block is the same in both
    assert np.allclose(H[2:, 2:], Href[2:, 2:], rtol=1e-12)


def test_hessian_transform_matches_oracle_restatement():
    """Parameters.transform_hessian (parameters.py:334-357) on log and bounds transforms."""
    values, act, tr = objective_trees("J2", True)
    P, Po = Parameters(values, act, tr), co.OracleParameters(*objective_trees("J2", True))
    x = np.array([0.2, -0.3, 0.1])
    P.set_active_values_from_flat(x, True); Po.set_active_values_from_flat(x, True)
    rng = np.random.default_rng(0)
    H = rng.standard_normal((3, 3)); H = H + H.T
    g = rng.standard_normal(3)
    H1, H2 = H.copy(), H.copy()
    P.transform_hessian(H1, g); co.transform_hessian(Po, H2, g)
    assert np.allclose(H1, H2, rtol=1e-15) and np.array_equal(H1, H1.T)
    assert not np.allclose(H1, H)


# ------------------------------------------------------------------------------------------ #
#  GPU                                                                                       #
# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_hessian_vs_reference(cuda_device, case):
    from cmad_b200.objectives import (Calibration, HessianResult, MPAdjointObjective, MPDirectAdjointObjective,
                                      SmallElasticPlastic)
    kind, mode = case.split(".")
    x = HS[f"{case}.x_canonical"]
    F, data, w = HS[f"{case}.F"], HS[f"{case}.data"], HS[f"{case}.weight"]
    P = Parameters(*objective_trees(kind, mode == "scaled"))
    obj = MPDirectAdjointObjective(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device,
                                   reference_qoi_cross_terms=True)
    r = obj.evaluate(x)
    assert isinstance(r, HessianResult)
    assert abs(r.J - HS[f"{case}.J"]) < 1e-11 * abs(r.J)
    assert np.abs(r.grad - HS[f"{case}.grad"]).max() < 1e-9 * np.abs(r.grad).max()
    assert hess_err(r.hessian, HS[f"{case}.hessian"]) < 1e-8, hess_err(r.hessian, HS[f"{case}.hessian"])
    assert np.array_equal(r.hessian, r.hessian.T)
    # same J and gradient as the plain adjoint objective; bit-reproducible
    P2 = Parameters(*objective_trees(kind, mode == "scaled"))
    ra = MPAdjointObjective(Calibration(SmallElasticPlastic(P2), data, w), F, device=cuda_device).evaluate(x)
    assert ra.J == r.J and np.array_equal(ra.grad, r.grad)
    r2 = obj.evaluate(x)
    assert np.array_equal(r.hessian, r2.hessian)
    # the complete Hessian differs from the reference's exactly when elastic parameters are active
    P3 = Parameters(*objective_trees(kind, mode == "scaled"))
    rc = MPDirectAdjointObjective(Calibration(SmallElasticPlastic(P3), data, w), F, device=cuda_device).evaluate(x)
    if mode == "scaled":
        assert hess_err(rc.hessian, r.hessian) < 1e-12
    else:
        assert hess_err(rc.hessian, r.hessian) > 1e-3
        assert hess_err(rc.hessian[2:, 2:], r.hessian[2:, 2:]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
def test_cuda_hessian_batch_vs_oracle_and_fd(cuda_device, kind):
    """A batch of points: (i) the summed Hessian of the first few points against the torch
    oracle run point by point; (ii) every Hessian column of the whole batch against central
    differences of the CUDA adjoint gradient in native parameter values."""
    import torch
    from cmad_b200.objectives import SmallElasticPlastic, gpu_local_evaluator
    from tests.helpers import param_tree
    from tests.test_objectives_host import _problem
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    active = ("E", "nu", "D", "S", "Y") + (("F", "L", "N") if kind == "hill" else ())
    values, act, tr = param_tree(kind, ("voce",), hill=hill, active=active)
    P = Parameters(values, act, tr)
    na = P.num_active_params
    model = SmallElasticPlastic(P)
    # (i) 4 points against the oracle
    sh, data, w = _problem(n=4, N=10, seed=3, kind=kind)
    out = gpu_local_evaluator(model, sh, data, w, "direct_adjoint", cuda_device)().cpu().numpy()
    Jo, go, Ho = 0.0, np.zeros(na), np.zeros((na, na))
    for p in range(4):
        F = np.repeat(np.eye(3)[:, :, None], 11, axis=2)
        for t in range(11):
            e = sh[t, :, p]
            F[:, :, t] += np.array([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]])
        Po = co.OracleParameters(*param_tree(kind, ("voce",), hill=hill, active=active))
        J, g, H = co.mp_objective_direct_adjoint(Po, F, data[:, :, p].T.reshape(3, 3, 11), w, co.ModelSpec())
        Jo += J; go += g; Ho += H                       # no transforms in this tree: native == canonical
    assert abs(out[0] - Jo) < 1e-11 * abs(Jo)
    assert np.abs(out[1:1 + na] - go).max() < 1e-9 * np.abs(go).max()
    Hg = out[1 + na:].reshape(na, na)
    assert hess_err(Hg, Ho) < 1e-8, hess_err(Hg, Ho)
    # (ii) FD of the adjoint gradient over a large batch
    sh, data, w = _problem(n=5000, N=10, seed=5, kind=kind)
    ev_h = gpu_local_evaluator(model, sh, data, w, "direct_adjoint", cuda_device)
    ev_g = gpu_local_evaluator(model, sh, data, w, "adjoint", cuda_device)
    out = ev_h().cpu().numpy()
    H = out[1 + na:].reshape(na, na)
    assert np.array_equal(H, H.T)
    assert np.array_equal(ev_h().cpu().numpy(), out)                   # deterministic
    p0 = P.flat_active_values(False).copy()
    for c in range(na):
        h = 1e-6 * abs(p0[c])
        gs = []
        for sgn in (1.0, -1.0):
            pp = p0.copy(); pp[c] += sgn * h
            P.set_active_values_from_flat(pp, False)
            gs.append(ev_g().cpu().numpy()[1:].copy())
        P.set_active_values_from_flat(p0, False)
        fd = (gs[0] - gs[1]) / (2 * h)
        d = np.sqrt(np.abs(np.diag(H)))
        assert (np.abs(fd - H[:, c]) / (d * d[c])).max() < 2e-5, (c, fd, H[:, c])


# ------------------------------------------------------------------------------------------ #
#  world size 2 (gloo): the Hessian objective's single all-reduce of 1 + P_a + P_a^2 doubles  #
# ------------------------------------------------------------------------------------------ #
def _hess_local_evaluator(P, Fs, datas, w):
    """Oracle-backed stand-in for the CUDA evaluator: this rank's (J, grad, H) in native
    parameter values, summed over its points."""
    import torch

    def ev():
        na = P.num_active_params
        out = np.zeros(1 + na + na * na)
        for F, data in zip(Fs, datas):
            Po = co.OracleParameters(*objective_trees("J2", True))
            Po.set_active_values_from_flat(P.flat_active_values(False), False)
            Po._flat_active_transforms = [None] * na              # native values, no chain rule here
            J, g, H = co.mp_objective_direct_adjoint(Po, F, data, w, co.ModelSpec())
            out += np.concatenate([[J], g, H.reshape(-1)])
        return torch.tensor(out)
    return ev


def _gloo_hess_worker(rank, world, port, Fs, datas, w, x, ret):
    import torch.distributed as dist
    from cmad_b200.objectives import BatchedMPObjective, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = Parameters(*objective_trees("J2", True))
    lo, hi = shard_range(len(Fs), rank, world)
    from cmad_b200.comm import WORLD
    res = BatchedMPObjective(P, _hess_local_evaluator(P, Fs[lo:hi], datas[lo:hi], w), group=WORLD).evaluate(x)
    if rank == 0:
        ret["J"], ret["grad"], ret["hessian"] = res.J, res.grad, res.hessian
    dist.destroy_process_group()


def test_two_rank_gloo_hessian_equals_single_process():
    import torch.multiprocessing as tmp
    from cmad_b200.objectives import BatchedMPObjective, HessianResult
    case = "J2.scaled"
    F, data, w, x = (HS[f"{case}.{k}"] for k in ("F", "data", "weight", "x_canonical"))
    F, data = F[:, :, :5], data[:, :, :5]
    Fs = [F, np.eye(3)[:, :, None] + 0.8 * (F - np.eye(3)[:, :, None]), F]
    datas = [data, 0.9 * data, 1.05 * data]                       # 3 points: ragged 2 + 1 split
    P = Parameters(*objective_trees("J2", True))
    single = BatchedMPObjective(P, _hess_local_evaluator(P, Fs, datas, w)).evaluate(x)
    assert isinstance(single, HessianResult)
    mgr = tmp.Manager(); ret = mgr.dict()
    port = 29000 + (os.getpid() + 7) % 2000
    tmp.spawn(_gloo_hess_worker, args=(2, port, Fs, datas, w, x, ret), nprocs=2, join=True)
    assert abs(ret["J"] - single.J) < 1e-12 * abs(single.J)
    assert np.allclose(ret["grad"], single.grad, rtol=1e-11, atol=0)
    assert hess_err(ret["hessian"], single.hessian) < 1e-11
    # canonical chain rule (log / bounds transforms) applied once, after the reduction: the
    # one-point oracle with transforms gives the same numbers
    Po = co.OracleParameters(*objective_trees("J2", True))
    J1, g1, H1 = co.mp_objective_direct_adjoint(Po, Fs[0], datas[0], w, co.ModelSpec(), x, True)
    P1 = Parameters(*objective_trees("J2", True))
    one = BatchedMPObjective(P1, _hess_local_evaluator(P1, Fs[:1], datas[:1], w)).evaluate(x)
    assert np.allclose(one.grad, g1, rtol=1e-11) and hess_err(one.hessian, H1) < 1e-11


# ------------------------------------------------------------------------------------------ #
#  PLANE_STRESS / UNIAXIAL_STRESS (KA5: the reference's Hessian check lives in plane stress)  #
# ------------------------------------------------------------------------------------------ #
_HDT_PATH = os.path.join(G, "ref_mp_hessian_dt.npz")
HDT = np.load(_HDT_PATH) if os.path.exists(_HDT_PATH) else None
CASES_DT = sorted({k.rsplit(".", 1)[0] for k in HDT.files}) if HDT is not None else []


@pytest.mark.parametrize("case", CASES_DT)
def test_torch_oracle_hessian_vs_reference_def_types(case):
    kind, mode, dtn = case.split(".")
    P = co.OracleParameters(*objective_trees(kind, mode == "scaled"))
    spec = co.ModelSpec(def_type=getattr(co, dtn))
    J, g, H = co.mp_objective_direct_adjoint(P, HDT[f"{case}.F"], HDT[f"{case}.data"], HDT[f"{case}.weight"],
                                             spec, HDT[f"{case}.x_canonical"], True,
                                             reference_qoi_cross_terms=True)
    assert abs(J - HDT[f"{case}.J"]) < 1e-11 * abs(J)
    assert np.abs(g - HDT[f"{case}.grad"]).max() < 1e-9 * np.abs(g).max()
    assert hess_err(H, HDT[f"{case}.hessian"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_DT)
def test_cuda_hessian_vs_reference_def_types(cuda_device, case):
    from cmad_b200 import objectives as ob
    kind, mode, dtn = case.split(".")
    x = HDT[f"{case}.x_canonical"]
    F, data, w = HDT[f"{case}.F"], HDT[f"{case}.data"], HDT[f"{case}.weight"]

    def run(compat):
        P = Parameters(*objective_trees(kind, mode == "scaled"))
        model = ob.SmallElasticPlastic(P, def_type=getattr(ob, dtn))
        return ob.MPDirectAdjointObjective(ob.Calibration(model, data, w), F, device=cuda_device,
                                           reference_qoi_cross_terms=compat).evaluate(x), P
    r, P = run(True)
    assert abs(r.J - HDT[f"{case}.J"]) < 1e-11 * abs(r.J)
    assert np.abs(r.grad - HDT[f"{case}.grad"]).max() < 1e-9 * np.abs(r.grad).max()
    assert hess_err(r.hessian, HDT[f"{case}.hessian"]) < 1e-8, hess_err(r.hessian, HDT[f"{case}.hessian"])
    # the complete Hessian = derivative of the CUDA adjoint gradient (canonical coordinates)
    rc, P = run(False)
    model = ob.SmallElasticPlastic(P, def_type=getattr(ob, dtn))
    grad_obj = ob.MPAdjointObjective(ob.Calibration(model, data, w), F, device=cuda_device)
    fd = np.zeros_like(rc.hessian)
    for c in range(len(x)):
        h = 1e-6 * max(abs(x[c]), 1e-2)
        xp_, xm_ = x.copy(), x.copy()
        xp_[c] += h; xm_[c] -= h
        fd[:, c] = (grad_obj.evaluate(xp_).grad - grad_obj.evaluate(xm_).grad) / (2 * h)
    # compared in parameter-scaled form H_ij x_i x_j (uniaxial stress does not see nu at all: its
    # row is zero up to rounding, so a sqrt(H_ii H_jj) normalisation would be meaningless there)
    sx = np.outer(np.abs(x), np.abs(x))
    assert np.abs((rc.hessian - fd) * sx).max() < 2e-5 * np.abs(rc.hessian * sx).max(), (rc.hessian, fd)
    if mode == "scaled":
        assert hess_err(rc.hessian, r.hessian) < 1e-12


@pytest.mark.gpu
def test_hessian_entry_point_edge_cases(cuda_device):
    """Argument errors and degenerate sizes of cmadx_mp_objective_hessian: no active parameter
    (J only), an active rotation-matrix entry (unsupported -> NotImplementedError, never a silent fallback), an
    unknown flag bit (EINVAL), and an elastic-only history (zero plastic sensitivity terms)."""
    import ctypes as C
    import torch
    from cmad_b200 import _lib as L
    from cmad_b200.objectives import SmallElasticPlastic, gpu_local_evaluator
    from tests.helpers import param_tree, rotation_matrix
    from tests.test_objectives_host import _problem
    sh, data, w = _problem(n=64, N=6, seed=9)
    # (1) nothing active
    v, a, t = param_tree("J2", ("voce",), active=())
    out = gpu_local_evaluator(SmallElasticPlastic(Parameters(v, a, t)), sh, data, w, "direct_adjoint",
                              cuda_device)().cpu().numpy()
    assert out.shape == (1,) and np.isfinite(out[0]) and out[0] > 0
    # (2) rotated axes with an active rotation-matrix entry: not differentiated twice -> refused
    v, a, t = param_tree("J2", ("voce",), rotation=rotation_matrix([1.0, 2.0, -0.5], 0.7))
    a["rotation matrix"] = True
    with pytest.raises(NotImplementedError):
        gpu_local_evaluator(SmallElasticPlastic(Parameters(v, a, t)), sh, data, w, "direct_adjoint", cuda_device)()
    # (3) elastic-only history: H = d2J/dp2 through the elastic constants only; flow-stress block zero
    v, a, t = param_tree("J2", ("voce",))
    P = Parameters(v, a, t)
    out = gpu_local_evaluator(SmallElasticPlastic(P), 1e-3 * sh, data, w, "direct_adjoint", cuda_device)().cpu().numpy()
    na = P.num_active_params
    H = out[1 + na:].reshape(na, na)
    assert np.all(H[2:, :] == 0.0) and np.all(H[:, 2:] == 0.0) and H[0, 0] > 0.0
    # (4) unknown flag bit straight through the C-ABI
    h = L.MpHistory()
    h.n, h.ld, h.nsteps, h.strain_comps = 0, 1, 1, 6
    res = torch.zeros(8, dtype=torch.float64, device=cuda_device)
    h.result, h.workspace = res.data_ptr(), res.data_ptr()
    mat = SmallElasticPlastic(P).material()
    rc = L.lib().cmadx_mp_objective_hessian(C.byref(mat), None, 0, C.byref(h), C.c_int32(2), None)
    assert rc == L.EINVAL


# ------------------------------------------------------------------------------------------ #
#  The reference's OTHER Hessian strategy: MPJVPObjective = jax.hessian of the whole loop     #
# ------------------------------------------------------------------------------------------ #
_HJ_PATH = os.path.join(G, "ref_mp_hessian_jvp.npz")
HJ = np.load(_HJ_PATH) if os.path.exists(_HJ_PATH) else None
CASES_JVP = sorted({k.rsplit(".", 1)[0] for k in HJ.files}) if HJ is not None else []


@pytest.mark.parametrize("case", CASES_JVP)
def test_complete_hessian_equals_the_references_jvp_strategy(case):
    """`MPJVPObjective` (cmad/objectives/mp_jvp_objective.py) differentiates the traced time loop
    with jax.hessian - no hand-assembled blocks.  Executed from the reference's own source, it
    agrees with the oracle's COMPLETE Hessian for every parameter set, and with the
    reference's `MPDirectAdjointObjective` only while no elastic parameter is active: the two
    strategies of the reference itself disagree on the d2J/dxi dp block (qoi.py:53-55)."""
    kind, mode = case.split(".")
    mk = lambda: co.OracleParameters(*objective_trees(kind, mode == "scaled"))
    args = (HJ[f"{case}.F"], HJ[f"{case}.data"], HJ[f"{case}.weight"], co.ModelSpec(), HJ[f"{case}.x_canonical"], True)
    J, g, H = co.mp_objective_direct_adjoint(mk(), *args)
    assert abs(J - HJ[f"{case}.J"]) < 1e-11 * abs(J)
    assert np.abs(g - HJ[f"{case}.grad"]).max() < 1e-9 * np.abs(g).max()
    assert hess_err(H, HJ[f"{case}.hessian"]) < 1e-8
    _, _, Hc = co.mp_objective_direct_adjoint(mk(), *args, reference_qoi_cross_terms=True)
    if mode == "scaled":
        assert hess_err(Hc, HJ[f"{case}.hessian"]) < 1e-8
    else:
        assert hess_err(Hc, HJ[f"{case}.hessian"]) > 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_JVP)
def test_cuda_jvp_strategy_vs_reference(cuda_device, case):
    """objectives.MPJVPObjective (traced Newton + adjoint + complete K2-H) against the reference's
    own MPJVPObjective run: J 1e-11, gradient 1e-9, Hessian 1e-8; the default
    MPDirectAdjointObjective (imperative Newton) gives the same Hessian."""
    from cmad_b200.objectives import Calibration, MPDirectAdjointObjective, MPJVPObjective, SmallElasticPlastic
    kind, mode = case.split(".")
    x = HJ[f"{case}.x_canonical"]
    F, data, w = HJ[f"{case}.F"], HJ[f"{case}.data"], HJ[f"{case}.weight"]
    P = Parameters(*objective_trees(kind, mode == "scaled"))
    obj = MPJVPObjective(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device)
    J, g = obj.evaluate_objective_and_grad(x)
    H = obj.evaluate_hessian(x)
    assert abs(obj.evaluate_objective(x) - J) == 0.0
    assert abs(J - HJ[f"{case}.J"]) < 1e-11 * abs(J)
    assert np.abs(g - HJ[f"{case}.grad"]).max() < 1e-9 * np.abs(g).max()
    assert hess_err(H, HJ[f"{case}.hessian"]) < 1e-8, hess_err(H, HJ[f"{case}.hessian"])
    P2 = Parameters(*objective_trees(kind, mode == "scaled"))
    r = MPDirectAdjointObjective(Calibration(SmallElasticPlastic(P2), data, w), F, device=cuda_device).evaluate(x)
    assert hess_err(r.hessian, HJ[f"{case}.hessian"]) < 1e-8


# ------------------------------------------------------------------------------------------ #
#  Rotated material axes in the direct-adjoint Hessian (anisotropic Hill, Voce + linear),     #
#  fixture ref_mp_hessian_rot.npz from the reference's MPDirectAdjointObjective                #
# ------------------------------------------------------------------------------------------ #
_HR_PATH = os.path.join(G, "ref_mp_hessian_rot.npz")
HR = np.load(_HR_PATH)
CASES_ROT = sorted({k.rsplit(".", 1)[0] for k in HR.files})


@pytest.mark.parametrize("case", CASES_ROT)
def test_torch_oracle_hessian_rotated_axes_vs_reference(case):
    kind, mode = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    assert np.abs(np.asarray(values["rotation matrix"]) - np.eye(3)).max() > 0.1
    P = co.OracleParameters(values, act, tr)
    F, data = HR[f"{case}.F"][:, :, :6], HR[f"{case}.data"][:, :, :6]        # a prefix keeps the torch run short ...
    J, g, H = co.mp_objective_direct_adjoint(P, F, data, HR[f"{case}.weight"], co.ModelSpec(),
                                             HR[f"{case}.x_canonical"], True, reference_qoi_cross_terms=True)
    assert np.isfinite(H).all() and np.array_equal(H, H.T)                   # ... the full history is compared on the GPU


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_ROT)
def test_cuda_hessian_rotated_axes_vs_reference(cuda_device, case):
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectAdjointObjective, SmallElasticPlastic
    kind, mode = case.split(".")
    x = HR[f"{case}.x_canonical"]
    F, data, w = HR[f"{case}.F"], HR[f"{case}.data"], HR[f"{case}.weight"]

    def run(compat):
        P = Parameters(*objective_trees(kind, mode == "scaled"))
        return MPDirectAdjointObjective(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device,
                                        reference_qoi_cross_terms=compat).evaluate(x), P
    r, _ = run(True)
    assert abs(r.J - HR[f"{case}.J"]) < 1e-11 * abs(r.J)
    assert np.abs(r.grad - HR[f"{case}.grad"]).max() < 1e-9 * np.abs(r.grad).max()
    assert hess_err(r.hessian, HR[f"{case}.hessian"]) < 1e-8, hess_err(r.hessian, HR[f"{case}.hessian"])
    # the complete Hessian = derivative of the CUDA adjoint gradient
    rc, P = run(False)
    grad_obj = MPAdjointObjective(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device)
    fd = np.zeros_like(rc.hessian)
    for c in range(len(x)):
        h = 1e-6 * max(abs(x[c]), 1e-2)
        xp_, xm_ = x.copy(), x.copy()
        xp_[c] += h; xm_[c] -= h
        fd[:, c] = (grad_obj.evaluate(xp_).grad - grad_obj.evaluate(xm_).grad) / (2 * h)
    assert hess_err(rc.hessian, 0.5 * (fd + fd.T)) < 2e-5, hess_err(rc.hessian, 0.5 * (fd + fd.T))


# ------------------------------------------------------------------------------------------ #
#  SmallRateElasticPlastic (FULL_3D, identity and rotated material axes): the reference's own   #
#  MPDirectAdjointObjective on the rate form (ref_mp_hessian_rate.npz).  The QoI reads the      #
#  state's stress, so the reference's Hessian is complete for this model.                       #
# ------------------------------------------------------------------------------------------ #
HRT = np.load(os.path.join(G, "ref_mp_hessian_rate.npz"))
CASES_RATE = sorted({k.rsplit(".", 1)[0] for k in HRT.files})


def test_rate_hessian_fixture_set():
    assert {c.split(".")[0] for c in CASES_RATE} == {"J2", "hill", "hosford", "hill_rot"}
    for case in CASES_RATE:
        H = HRT[f"{case}.hessian"]
        assert np.isfinite(H).all() and np.abs(H - H.T).max() < 1e-8 * np.abs(H).max()


@pytest.mark.parametrize("case", ["J2.scaled", "hill_rot.native"])
def test_torch_oracle_rate_hessian_vs_reference(case):
    kind, mode = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    J, g, H = co.mp_objective_direct_adjoint(P, HRT[f"{case}.F"], HRT[f"{case}.data"], HRT[f"{case}.weight"], spec,
                                             HRT[f"{case}.x_canonical"], True, reference_qoi_cross_terms=True)
    assert abs(J - HRT[f"{case}.J"]) < 1e-11 * abs(J)
    assert np.abs(g - HRT[f"{case}.grad"]).max() < 1e-9 * np.abs(g).max()
    assert hess_err(H, HRT[f"{case}.hessian"]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_RATE)
def test_cuda_rate_hessian_vs_reference(cuda_device, case):
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectAdjointObjective, SmallRateElasticPlastic
    kind, mode = case.split(".")
    x = HRT[f"{case}.x_canonical"]
    F, data, w = HRT[f"{case}.F"], HRT[f"{case}.data"], HRT[f"{case}.weight"]

    def run(compat):
        P = Parameters(*objective_trees(kind, mode == "scaled"))
        assert np.array_equal(P.active_idx, HRT[f"{case}.active_idx"])
        return MPDirectAdjointObjective(Calibration(SmallRateElasticPlastic(P), data, w), F, device=cuda_device,
                                        reference_qoi_cross_terms=compat).evaluate(x), P
    r, _ = run(True)
    assert abs(r.J - HRT[f"{case}.J"]) < 1e-11 * abs(r.J)
    assert np.abs(r.grad - HRT[f"{case}.grad"]).max() < 1e-9 * np.abs(r.grad).max()
    assert hess_err(r.hessian, HRT[f"{case}.hessian"]) < 1e-8, hess_err(r.hessian, HRT[f"{case}.hessian"])
    assert np.array_equal(r.hessian, r.hessian.T)
    # complete = reference-compatible for this model (no parameter enters the QoI)
    rc, P = run(False)
    assert np.array_equal(rc.hessian, r.hessian)
    # ... and it is the derivative of the CUDA adjoint gradient once every step's Newton solve has
    # converged: the reference's 10 imperative iterations leave step 6 of `hill.native` unconverged (it
    # needs 13; the fixture - and the comparison above - carry that state), so the check runs with 60
    from cmad_b200 import NewtonSettings
    from cmad_b200.objectives import _single_point_objective
    nw = NewtonSettings(mode="imperative", max_iters=60, abs_tol=1e-14, rel_tol=1e-14)
    qoi = Calibration(SmallRateElasticPlastic(P), data, w)
    Hc = _single_point_objective(qoi, F, "direct_adjoint", cuda_device, newton=nw).evaluate(x).hessian
    grad_obj = _single_point_objective(qoi, F, "adjoint", cuda_device, newton=nw)
    fd = np.zeros_like(Hc)
    for c in range(len(x)):
        h = 1e-6 * max(abs(x[c]), 1e-2)
        xp_, xm_ = x.copy(), x.copy()
        xp_[c] += h; xm_[c] -= h
        fd[:, c] = (grad_obj.evaluate(xp_).grad - grad_obj.evaluate(xm_).grad) / (2 * h)
    assert hess_err(Hc, 0.5 * (fd + fd.T)) < 2e-5, hess_err(Hc, 0.5 * (fd + fd.T))


# ------------------------------------------------------------------------------------------ #
#  The rate form under PLANE_STRESS / UNIAXIAL_STRESS (n_xi 8 / 12), identity and rotated axes  #
#  (ref_mp_hessian_rate_dt.npz, the reference's own MPDirectAdjointObjective)                   #
# ------------------------------------------------------------------------------------------ #
HRD = np.load(os.path.join(G, "ref_mp_hessian_rate_dt.npz"))
CASES_RATE_DT = sorted({k.rsplit(".", 1)[0] for k in HRD.files})


def hess_err_sensitive(H, Href, grad, x):
    """`hess_err` over the parameters the objective depends on.  Under uniaxial stress J does not see
    nu in some settings (gradient entry 1e-10 of the others AND a Hessian diagonal of rounding noise):
    such rows are required to vanish on the scale of the others instead of being compared entry by
    entry against noise.  (A parameter with a vanishing gradient but a finite reference diagonal - the
    elastic rows of the reference's incomplete Hessian - is compared like any other.)"""
    s = np.abs(grad * x)
    dd = np.abs(np.diag(Href) * x * x)
    keep = (s > 1e-9 * s.max()) | (dd > 1e-9 * dd.max())
    scale = dd[keep].max()
    for i in np.nonzero(~keep)[0]:
        assert (np.abs(H[i] * x[i] * x).max() < 1e-6 * scale) and (np.abs(Href[i] * x[i] * x).max() < 1e-6 * scale)
    return hess_err(H[np.ix_(keep, keep)], Href[np.ix_(keep, keep)])


def test_rate_def_type_hessian_fixture_set():
    assert {c.rsplit(".", 1)[1] for c in CASES_RATE_DT} == {"PLANE_STRESS", "UNIAXIAL_STRESS"}
    assert any(c.startswith("hill_rot") for c in CASES_RATE_DT)


@pytest.mark.parametrize("case", [c for c in CASES_RATE_DT if c.startswith(("J2.scaled.PLANE", "hill_rot.native.UNIAXIAL"))])
def test_torch_oracle_rate_hessian_vs_reference_def_types(case):
    kind, mode, dtn = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec(kind="small_rate_elastic_plastic", def_type=getattr(co, dtn))
    J, g, H = co.mp_objective_direct_adjoint(P, HRD[f"{case}.F"], HRD[f"{case}.data"], HRD[f"{case}.weight"], spec,
                                             HRD[f"{case}.x_canonical"], True, reference_qoi_cross_terms=True)
    assert abs(J - HRD[f"{case}.J"]) < 1e-10 * abs(J)
    assert np.abs(g - HRD[f"{case}.grad"]).max() < 1e-8 * np.abs(g).max()
    xn = HRD[f"{case}.active_native"] if mode == "native" else np.ones_like(g)
    assert hess_err_sensitive(H, HRD[f"{case}.hessian"], HRD[f"{case}.grad"], xn) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_RATE_DT)
def test_cuda_rate_hessian_vs_reference_def_types(cuda_device, case):
    from cmad_b200 import NewtonSettings
    from cmad_b200 import objectives as ob
    kind, mode, dtn = case.split(".")
    x = HRD[f"{case}.x_canonical"]
    F, data, w = HRD[f"{case}.F"], HRD[f"{case}.data"], HRD[f"{case}.weight"]
    P = Parameters(*objective_trees(kind, mode == "scaled"))
    assert np.array_equal(P.active_idx, HRD[f"{case}.active_idx"])
    qoi = ob.Calibration(ob.SmallRateElasticPlastic(P, def_type=getattr(ob, dtn)), data, w)
    r = ob.MPDirectAdjointObjective(qoi, F, device=cuda_device, reference_qoi_cross_terms=True).evaluate(x)
    assert abs(r.J - HRD[f"{case}.J"]) < 1e-10 * abs(r.J)
    assert np.abs(r.grad - HRD[f"{case}.grad"]).max() < 1e-8 * np.abs(r.grad).max()
    xn = HRD[f"{case}.active_native"] if mode == "native" else np.ones_like(x)
    assert hess_err_sensitive(r.hessian, HRD[f"{case}.hessian"], HRD[f"{case}.grad"], xn) < 1e-7
    # with every Newton solve converged: the derivative of the CUDA adjoint gradient
    nw = NewtonSettings(mode="imperative", max_iters=60, abs_tol=1e-14, rel_tol=1e-14)
    Hc = ob._single_point_objective(qoi, F, "direct_adjoint", cuda_device, newton=nw).evaluate(x).hessian
    grad_obj = ob._single_point_objective(qoi, F, "adjoint", cuda_device, newton=nw)
    fd = np.zeros_like(Hc)
    for c in range(len(x)):
        h = 1e-6 * max(abs(x[c]), 1e-2)
        xp_, xm_ = x.copy(), x.copy()
        xp_[c] += h; xm_[c] -= h
        fd[:, c] = (grad_obj.evaluate(xp_).grad - grad_obj.evaluate(xm_).grad) / (2 * h)
    assert hess_err_sensitive(Hc, 0.5 * (fd + fd.T), HRD[f"{case}.grad"], xn) < 2e-5


# ------------------------------------------------------------------------------------------ #
#  SmallElasticPlastic under PLANE_STRESS / UNIAXIAL_STRESS with ROTATED material axes          #
#  (mp_hess_dt_kernel<.., ROT>; ref_mp_hessian_dt_rot.npz from the reference's own run)          #
# ------------------------------------------------------------------------------------------ #
HDR = np.load(os.path.join(G, "ref_mp_hessian_dt_rot.npz"))
CASES_DT_ROT = sorted({k.rsplit(".", 1)[0] for k in HDR.files})


def test_rotated_def_type_hessian_fixture_set():
    assert {c.rsplit(".", 1)[1] for c in CASES_DT_ROT} == {"PLANE_STRESS", "UNIAXIAL_STRESS"}
    assert all(c.startswith("hill_rot") for c in CASES_DT_ROT)


@pytest.mark.parametrize("case", [c for c in CASES_DT_ROT if c.startswith("hill_rot.native")])
def test_torch_oracle_hessian_rotated_def_types_vs_reference(case):
    kind, mode, dtn = case.split(".")
    P = co.OracleParameters(*objective_trees(kind, mode == "scaled"))
    spec = co.ModelSpec(def_type=getattr(co, dtn))
    J, g, H = co.mp_objective_direct_adjoint(P, HDR[f"{case}.F"], HDR[f"{case}.data"], HDR[f"{case}.weight"], spec,
                                             HDR[f"{case}.x_canonical"], True, reference_qoi_cross_terms=True)
    xn = HDR[f"{case}.active_native"]
    assert abs(J - HDR[f"{case}.J"]) < 1e-10 * abs(J)
    assert hess_err_sensitive(H, HDR[f"{case}.hessian"], HDR[f"{case}.grad"], xn) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_DT_ROT)
def test_cuda_hessian_rotated_def_types_vs_reference(cuda_device, case):
    from cmad_b200 import NewtonSettings
    from cmad_b200 import objectives as ob
    kind, mode, dtn = case.split(".")
    x = HDR[f"{case}.x_canonical"]
    F, data, w = HDR[f"{case}.F"], HDR[f"{case}.data"], HDR[f"{case}.weight"]
    P = Parameters(*objective_trees(kind, mode == "scaled"))
    assert np.array_equal(P.active_idx, HDR[f"{case}.active_idx"])
    qoi = ob.Calibration(ob.SmallElasticPlastic(P, def_type=getattr(ob, dtn)), data, w)
    r = ob.MPDirectAdjointObjective(qoi, F, device=cuda_device, reference_qoi_cross_terms=True).evaluate(x)
    xn = HDR[f"{case}.active_native"] if mode == "native" else np.ones_like(x)
    assert abs(r.J - HDR[f"{case}.J"]) < 1e-10 * abs(r.J)
    gs = np.abs(HDR[f"{case}.grad"] * xn)
    keep = gs > 1e-9 * gs.max()
    assert np.abs(r.grad - HDR[f"{case}.grad"])[keep].max() < 1e-8 * np.abs(r.grad).max()
    assert hess_err_sensitive(r.hessian, HDR[f"{case}.hessian"], HDR[f"{case}.grad"], xn) < 1e-7
    # the complete Hessian, every Newton solve converged: the derivative of the CUDA adjoint gradient
    nw = NewtonSettings(mode="imperative", max_iters=60, abs_tol=1e-14, rel_tol=1e-14)
    Hc = ob._single_point_objective(qoi, F, "direct_adjoint", cuda_device, newton=nw).evaluate(x).hessian
    grad_obj = ob._single_point_objective(qoi, F, "adjoint", cuda_device, newton=nw)
    fd = np.zeros_like(Hc)
    for c in range(len(x)):
        h = 1e-6 * max(abs(x[c]), 1e-2)
        xp_, xm_ = x.copy(), x.copy()
        xp_[c] += h; xm_[c] -= h
        fd[:, c] = (grad_obj.evaluate(xp_).grad - grad_obj.evaluate(xm_).grad) / (2 * h)
    assert hess_err_sensitive(Hc, 0.5 * (fd + fd.T), HDR[f"{case}.grad"], xn) < 2e-5
