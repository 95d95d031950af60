"""FE QoIs (cmad_b200/fe_qoi.py) and their gradients through the load steps: FEDisplacementMatch,
FELoadMatch (a QoI of the reactions: the residual's direct dependence on the parameters and on
xi_{n-1} enters the gradient), FEWeightedSum - direct == adjoint == central finite differences,
all over the oracle on CPU; the CUDA K6 kernels on GPU against the same numbers."""
import copy

import numpy as np
import pytest

from cmad_b200 import fe_driver as drv, fe_qoi
from tests.test_fe_driver import (_gradient_problem, oracle_assembler, oracle_jvp, oracle_vjp_pair)

TS = np.array([0.0, 0.4, 0.7, 1.0])
TIGHT = {"abs tol": 1e-13, "rel tol": 1e-13, "max iters": 15}


def _qoi(arr, nodes, bcs):
    """displacement match against a perturbed field + load match on the pulled face (component x)."""
    rng = np.random.default_rng(8)
    data_u = np.zeros((len(TS), nodes.shape[0], 3))
    for k, t in enumerate(TS):
        data_u[k, :, 0] = 0.0028 * t * nodes[:, 0]
        data_u[k] += 2e-5 * rng.standard_normal(data_u[k].shape)
    ramp = bcs.indices[np.isclose(nodes[bcs.indices // 3, 0], 1.0) & (bcs.indices % 3 == 0)]
    load = fe_qoi.FELoadMatch([ramp], TS, np.array([0.0, 150.0, 210.0, 230.0]), weight=1e-6)
    disp = fe_qoi.FEDisplacementMatch(arr, TS, data_u, weight=3.0)
    return fe_qoi.FEWeightedSum([disp, load]), load, disp


def test_qoi_units():
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, "hex8")
    q, load, disp = _qoi(arr, nodes, bcs)
    assert q.needs_residual and load.needs_residual and not disp.needs_residual
    rng = np.random.default_rng(0)
    U = 1e-3 * rng.standard_normal(arr.n_dofs); R = rng.standard_normal(arr.n_dofs)
    s = fe_qoi.StepState(U=U, t=0.7, t_prev=0.4, R=R, step=2)
    # dU / dR against finite differences of value
    g, gR = q.dU(s), q.dR(s)
    for vec, grad, attr in ((U, g, "U"), (R, gR, "R")):
        d = rng.standard_normal(vec.shape)
        h = 1e-6 * np.abs(vec).max()
        sp, sm = copy.copy(s), copy.copy(s)
        setattr(sp, attr, vec + h * d); setattr(sm, attr, vec - h * d)
        fd = (q.value(sp) - q.value(sm)) / (2 * h)
        assert abs(fd - grad @ d) < 1e-6 * abs(fd) + 1e-18, (attr, fd, grad @ d)
    assert np.allclose(fe_qoi.reaction_series(load, [R, 2 * R]), [[R[load.eqs[0]].sum()], [2 * R[load.eqs[0]].sum()]])
    with pytest.raises(ValueError):
        fe_qoi.FEDisplacementMatch(arr, TS, np.zeros((2, nodes.shape[0], 3)))
    with pytest.raises(ValueError):
        fe_qoi.FELoadMatch([load.eqs[0]], TS, np.zeros((4, 2)))


def test_load_and_displacement_match_gradients_direct_adjoint_fd():
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, "tet4")
    q, load, _ = _qoi(arr, nodes, bcs)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    asm = oracle_assembler(values, arr, scatter, len(pattern.rows))
    Jd, gd = drv.fe_direct_gradient(asm, oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                    lambda c: z(), TS, 5, TIGHT, qoi=q)
    vjp, vjp_disp = oracle_vjp_pair(values, P, arr)
    Ja, ga = drv.fe_adjoint_gradient(asm, vjp, vjp_disp, pattern, bcs, np.zeros(arr.n_dofs), z(), TS, 5, TIGHT, qoi=q)
    assert abs(Ja - Jd) < 1e-14 * abs(Jd)
    assert np.abs(ga - gd).max() < 1e-8 * np.abs(gd).max(), (ga, gd)
    # the load term alone matters (its share of the gradient is not negligible)
    _, gl = drv.fe_direct_gradient(asm, oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                   lambda c: z(), TS, 5, TIGHT, qoi=load)
    assert np.abs(gl).max() > 1e-3 * np.abs(gd).max()

    def J_of(v):
        return drv.fe_quasistatic_drive(oracle_assembler(v, arr, scatter, len(pattern.rows)), pattern, bcs,
                                        np.zeros(arr.n_dofs), z(), TS, TIGHT, qoi=q)[2]
    assert abs(J_of(values) - Jd) < 1e-13 * abs(Jd)
    paths = [("elastic", "E"), ("elastic", "nu"), ("plastic", "flow stress", "hardening", "voce", "D"),
             ("plastic", "flow stress", "hardening", "voce", "S"), ("plastic", "flow stress", "initial yield", "Y")]
    for c, path in enumerate(paths):
        def bump(h):
            v = copy.deepcopy(values); d = v
            for k in path[:-1]:
                d = d[k]
            d[path[-1]] = d[path[-1]] * (1 + h)
            return J_of(v)
        d = values
        for k in path:
            d = d[k]
        h = 1e-5
        fd = (bump(h) - bump(-h)) / (2 * h * d)
        assert abs(fd - gd[c]) < 5e-5 * abs(gd[c]) + 1e-12 * np.abs(gd).max(), (path, fd, gd[c])


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["hex8", "tet4"])
def test_cuda_qoi_gradients_match_oracle(cuda_device, family):
    import torch
    from cmad_b200 import active_param_ids, fe, material_from_values
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, family)
    q, _, _ = _qoi(arr, nodes, bcs)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    asm_o = oracle_assembler(values, arr, scatter, len(pattern.rows))
    Jo, go = drv.fe_direct_gradient(asm_o, oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                    lambda c: z(), TS, 5, TIGHT, qoi=q)
    mat, pid = material_from_values(values), active_param_ids(P)
    arr_d = arr.to(cuda_device)
    r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    nw = fe.fe_newton_settings(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
    asm = drv.cuda_assembler(mat, nw, arr_d, r_plan, k_plan)
    zd = lambda: torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    vjp, vjp_disp = drv.cuda_vjp(mat, arr_d, pid)
    Ja, ga = drv.fe_adjoint_gradient(asm, vjp, vjp_disp, pattern, bcs, np.zeros(arr.n_dofs), zd(), TS, 5, TIGHT, qoi=q)
    Jd, gd = drv.fe_direct_gradient(asm, drv.cuda_jvp(mat, arr_d, r_plan, pid), pattern, bcs, np.zeros(arr.n_dofs),
                                    zd(), lambda c: zd(), TS, 5, TIGHT, qoi=q)
    assert abs(Ja - Jo) < 1e-10 * abs(Jo) and abs(Jd - Jo) < 1e-10 * abs(Jo)
    assert np.abs(ga - go).max() < 1e-8 * np.abs(go).max(), (ga, go)
    assert np.abs(gd - go).max() < 1e-8 * np.abs(go).max(), (gd, go)
