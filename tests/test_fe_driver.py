"""FE quasi-static driver (global Newton + embedded BCs + cubic line search + load
steps): BASELINE.json configs[0] (examples/elastic_plastic_uniaxial.yaml: unit cube,
J2+Voce, pulled in x to 3x yield strain in 5 steps, QoI fe_displacement_l2).

CPU: the driver over the oracle assembler reproduces the deck's own claim - every
point's terminal sigma_xx equals the analytical uniaxial J2+Voce flow stress, lateral
stresses vanish - plus unit tests of the cubic line search.  GPU: the same loop over
the CUDA assembler (K3 + K5) gives the same trajectory as over the oracle."""
import math

import numpy as np
import pytest

from cmad_b200 import fe_driver as drv, fe_mesh
from oracle import analytic, fe_oracle, oracle_c as oc

LOCAL_NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)


def uniaxial_cube(div, family="hex8"):
    """Mesh + BCs of examples/elastic_plastic_uniaxial.yaml:58-66 on a div^3 cube."""
    nodes, conn = fe_mesh.structured_hex_mesh((div,) * 3)
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr = fe_mesh.block_arrays(nodes, conn)
    nid = np.arange(nodes.shape[0])
    on = lambda ax, v: nid[np.isclose(nodes[:, ax], v)]
    pin = np.concatenate([on(0, 0.0) * 3 + 0, on(1, 0.0) * 3 + 1, on(2, 0.0) * 3 + 2])
    ramp = on(0, 1.0) * 3 + 0
    idx = np.concatenate([pin, ramp])
    bcs = drv.DirichletBCs(idx, lambda t: np.concatenate([np.zeros(len(pin)), np.full(len(ramp), 0.003 * t)]))
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy())
    pattern = drv.SparsePattern(ur, uc, arr.n_dofs)
    return nodes, arr, bcs, pattern, scatter


def oracle_assembler(values, arr, scatter, n_unique, record=None):
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **LOCAL_NEWTON)
    eq = arr.elem_eq.numpy(); geo = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())

    def assemble(U, xi_prev):
        o = fe_oracle.assemble_block(prob, eq, U, xi_prev, *geo)
        if record is not None:
            record["sigma"] = o["sigma"]
        return o["R"], fe_oracle.coo_dedup_sum(o["K_elem"].reshape(-1), scatter, n_unique), o["xi"]
    return assemble


def uniaxial_flow_stress(eps, E=200e3, Y=200.0, S=200.0, D=20.0):
    """sigma with eps = sigma/E + alpha, sigma = Y + S(1 - exp(-D alpha))."""
    a = 0.0
    for _ in range(60):
        g = Y + S * (1 - math.exp(-D * a)) - E * (eps - a)
        a -= g / (S * D * math.exp(-D * a) + E)
    return E * (eps - a), a


def test_cubic_min_and_line_search_units():
    """tests/util/test_line_search.py:15-163 (restated): exact minimiser of a cubic,
    fallback on a negative radicand, full-step acceptance, contraction clipping."""
    f = lambda a: a ** 3 + 0.9 * a ** 2 - 1.2 * a             # the reference test's cubic: minimum at 0.4
    df = lambda a: 3.0 * a ** 2 + 1.8 * a - 1.2
    assert abs(drv.cubic_min(f(0.0), df(0.0), 1.0, f(1.0), df(1.0)) - 0.4) < 1e-12
    # negative radicand (d1^2 < dphi_0 * slope_a): no real interior minimiser -> a / 2
    assert drv.cubic_min(0.0, -4.0, 1.0, -3.0, -4.0) == 0.5
    a, aux, n = drv.line_search(lambda al: (0.5 * (1 - al) ** 2, -(1 - al), al), 0.5, -1.0, None, None)
    assert (a, aux, n) == (1.0, 1.0, 1)
    phis = iter([10.0, 10.0, 10.0, 10.0])
    tried = []
    def bad(al):
        tried.append(al)
        return next(phis), 1.0, al
    a, aux, n = drv.line_search(bad, 0.5, -1.0, None, "base")
    assert n == 4 and a == tried[0] and all(0.5 * tried[k] <= tried[k + 1] <= 0.9 * tried[k] for k in range(3))
    a, _, n = drv.line_search(lambda al: (float("nan"), 0.0, al), 0.5, -1.0, {"max evals": 3}, "base")
    assert n == 3 and a == 1.0                                  # never finite: lowest-merit default


@pytest.mark.parametrize("family,div", [("hex8", 2), ("tet4", 2)])
def test_uniaxial_cube_reaches_analytic_flow_stress(family, div):
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, arr, bcs, pattern, scatter = uniaxial_cube(div, family)
    rec = {}
    asm = oracle_assembler(values, arr, scatter, len(pattern.rows), rec)
    ts = np.linspace(0.0, 1.0, 6)                               # num steps 5, step size 0.2
    wdet = (arr.det * arr.quad_w[None, :]).numpy()
    vol = wdet.sum()
    qoi = lambda U, t, tp: (t - tp) / (ts[-1] * vol) * drv.displacement_l2_step(arr.N.numpy(), wdet, arr.elem_eq.numpy(), U)
    U_steps, xi, J, logs = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs),
                                                    np.zeros((arr.n_elems, arr.n_ip, 7)), ts, None, qoi)
    assert all(l.iters <= 10 and l.residual_norms[-1] < 1e-8 for l in logs)
    sig_ref, alpha_ref = uniaxial_flow_stress(0.003)
    sig = rec["sigma"]                                          # at the last assembly = converged state
    assert np.abs(sig[:, :, 0] - sig_ref).max() < 1e-8 * sig_ref
    assert np.abs(sig[:, :, 1:]).max() < 1e-8 * sig_ref
    assert np.abs(xi[:, :, 6] - alpha_ref).max() < 1e-10
    # homogeneous field u_x = 0.003 x: J = sum_k dt * (1/V) int |u(t_k)|^2 (+ lateral contraction)
    assert J > 0 and np.isfinite(J)
    ux = U_steps[-1].reshape(-1, 3)[:, 0]
    assert np.abs(ux - 0.003 * nodes[:, 0]).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("family,div", [("hex8", 8), ("tet4", 4)])
def test_cuda_driver_matches_oracle_driver(cuda_device, family, div):
    """configs[0] itself (8x8x8 hexes = 4096 material points, 5 load steps) through the
    CUDA assembler vs through the oracle assembler: same Newton histories, same U, J."""
    import torch
    from cmad_b200 import fe, material_from_values
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, arr, bcs, pattern, scatter = uniaxial_cube(div, family)
    # a perturbed, non-homogeneous variant as well: body of the cube softened by a random field
    ts = np.linspace(0.0, 1.0, 6)
    wdet = (arr.det * arr.quad_w[None, :]).numpy()
    qoi = lambda U, t, tp: (t - tp) / wdet.sum() * drv.displacement_l2_step(arr.N.numpy(), wdet, arr.elem_eq.numpy(), U)
    arr_d = arr.to(cuda_device)
    r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    keep = {}
    asm = drv.cuda_assembler(material_from_values(values), fe.fe_newton_settings(**LOCAL_NEWTON), arr_d,
                             r_plan, k_plan, keep)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Ug, xig, Jg, lg = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs), xi0, ts, None, qoi)
    asm_o = oracle_assembler(values, arr, scatter, len(pattern.rows))
    Uo, xio, Jo, lo = drv.fe_quasistatic_drive(asm_o, pattern, bcs, np.zeros(arr.n_dofs),
                                               np.zeros((arr.n_elems, arr.n_ip, 7)), ts, None, qoi)
    assert [l.iters for l in lg] == [l.iters for l in lo]
    assert [l.assemblies for l in lg] == [l.assemblies for l in lo]
    assert np.abs(Ug - Uo).max() < 1e-10 * np.abs(Uo).max()
    assert abs(Jg - Jo) < 1e-10 * abs(Jo)
    assert np.abs(xig.cpu().numpy() - xio).max() < 1e-10 * np.abs(xio).max()
    sig_ref, _ = uniaxial_flow_stress(0.003)
    out = fe.fe_block_launch(material_from_values(values), fe.fe_newton_settings(**LOCAL_NEWTON), arr_d,
                             torch.from_numpy(Ug[-1]).to(cuda_device), xig, ("xi", "sigma"))
    # re-evaluation at the converged state (xi_prev = xi*: an elastic step that returns the same stress)
    assert float((out["sigma"][:, :, 0] - sig_ref).abs().max()) < 1e-7 * sig_ref


# ---------------------------------------------------------------- FE gradient (direct sensitivities)
def _gradient_problem(div=2, family="hex8"):
    from cmad_b200 import Parameters
    from tests.helpers import param_tree
    values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    nodes, arr, bcs, pattern, scatter = uniaxial_cube(div, family)
    # distort the interior so the fields are not homogeneous
    rng = np.random.default_rng(4)
    inner = np.all((nodes > 1e-9) & (nodes < 1 - 1e-9), axis=1)
    nodes = nodes.copy(); nodes[inner] += 0.08 / div * rng.uniform(-1, 1, size=(inner.sum(), 3))
    conn = fe_mesh.structured_hex_mesh((div,) * 3)[1]
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr = fe_mesh.block_arrays(nodes, conn)
    return values, P, nodes, arr, bcs, pattern, scatter


def oracle_jvp(values, P, arr):
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    eq = arr.elem_eq.numpy(); geo = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    na = len(prob_eval.active_pid)

    def jvp(U, xi_prev, xi_state, c, dxi_prev, dU):
        dp = np.zeros(na); dp[c] = 1.0
        o = fe_oracle.block_jvp(prob_eval, eq, U, xi_prev, xi_state, *geo, dp, dxi_prev, dU=dU)
        dR = np.zeros(arr.n_dofs); np.add.at(dR, eq.reshape(-1), o["R_elem"].reshape(-1))
        return dR, o["xi"]
    return jvp


def _qois(arr, ts):
    wdet = (arr.det * arr.quad_w[None, :]).numpy()
    N, eq = arr.N.numpy(), arr.elem_eq.numpy()
    c = 1.0 / (ts[-1] * wdet.sum())
    q = lambda U, t, tp: c * (t - tp) * drv.displacement_l2_step(N, wdet, eq, U)
    dq = lambda U, t, tp: c * (t - tp) * drv.displacement_l2_step_dU(N, wdet, eq, U)
    return q, dq


def test_direct_gradient_vs_central_fd_of_the_trajectory():
    """`cmad gradient` on the FE deck, restated: dJ/dp of the displacement-L2 QoI through
    3 load steps by forward sensitivities (K6 blocks + one factorisation per step), against
    central differences of J over re-solved trajectories (all over the oracle assembler)."""
    import copy
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem()
    # t = 1/3 would put the homogeneous part of the field exactly AT first yield (0.003 t =
    # Y/E), where J(p) has a kink in E and Y and a central difference averages the two
    # one-sided slopes; keep every step clear of it
    ts = np.array([0.0, 0.4, 0.7, 1.0])
    q, dq = _qois(arr, ts)
    tight = {"abs tol": 1e-13, "rel tol": 1e-13, "max iters": 15}
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    J, g = drv.fe_direct_gradient(oracle_assembler(values, arr, scatter, len(pattern.rows)),
                                  oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                  lambda c: z(), ts, 5, tight, q, dq)

    def J_of(v):
        return drv.fe_quasistatic_drive(oracle_assembler(v, arr, scatter, len(pattern.rows)), pattern, bcs,
                                        np.zeros(arr.n_dofs), z(), ts, tight, q)[2]
    assert abs(J_of(values) - J) < 1e-14 * abs(J)
    paths = [("elastic", "E"), ("elastic", "nu"), ("plastic", "flow stress", "hardening", "voce", "D"),
             ("plastic", "flow stress", "hardening", "voce", "S"), ("plastic", "flow stress", "initial yield", "Y")]
    assert np.abs(g).min() > 0
    for c, path in enumerate(paths):
        def bump(h):
            v = copy.deepcopy(values); d = v
            for k in path[:-1]:
                d = d[k]
            d[path[-1]] = d[path[-1]] * (1 + h)
            return J_of(v)
        h = 1e-5
        d = values
        for k in path:
            d = d[k]
        fd = (bump(h) - bump(-h)) / (2 * h * d)
        assert abs(fd - g[c]) < 2e-5 * abs(g[c]) + 1e-16, (path, fd, g[c])


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["hex8", "tet4"])
def test_cuda_direct_gradient_matches_oracle(cuda_device, family):
    import torch
    from cmad_b200 import active_param_ids, fe, material_from_values
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(3, family)
    ts = np.array([0.0, 0.4, 0.7, 1.0])
    q, dq = _qois(arr, ts)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    Jo, go = drv.fe_direct_gradient(oracle_assembler(values, arr, scatter, len(pattern.rows)),
                                    oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                    lambda c: z(), ts, 5, None, q, dq)
    arr_d = arr.to(cuda_device)
    mat = material_from_values(values)
    r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    zd = lambda: torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Jg, gg = drv.fe_direct_gradient(
        drv.cuda_assembler(mat, fe.fe_newton_settings(**LOCAL_NEWTON), arr_d, r_plan, k_plan),
        drv.cuda_jvp(mat, arr_d, r_plan, active_param_ids(P)), pattern, bcs, np.zeros(arr.n_dofs), zd(),
        lambda c: zd(), ts, 5, None, q, dq)
    assert abs(Jg - Jo) < 1e-10 * abs(Jo)
    assert np.abs(gg - go).max() < 1e-8 * np.abs(go).max(), (gg, go)
    # canonical-coordinate chain rule as `cmad gradient` reports it (parameters.py:326-331)
    gc = gg.copy(); P.transform_grad(gc)
    assert np.all(np.isfinite(gc))


# ------------------------------------------------ examples/notch_hosford.yaml at its native size
def _notch_problem():
    """examples/notch_hosford.yaml:18-55 on the reference's own mesh (fixture produced by
    tests/golden/make_notch_mesh.py): near-Tresca Hosford a=100, E=1000, nu=.25, Y=2,
    Voce S=10 D=2; local Newton 500 iters / 1e-12 with a 100-eval line search; global
    Newton 15 iters / 1e-8; symmetry on the three min faces, y-max face pulled 0.01 t;
    4 steps of size 1."""
    import os
    from tests.helpers import param_tree
    m = np.load(os.path.join(os.path.dirname(__file__), "golden", "notch_mesh.npz"))
    nodes, tets = m["nodes"], m["tets"]
    values, _, _ = param_tree("hosford", ("voce",), a=100.0, elastic={"E": 1000.0, "nu": 0.25}, active=())
    values["plastic"]["flow stress"]["initial yield"]["Y"] = 2.0
    values["plastic"]["flow stress"]["hardening"]["voce"] = {"S": 10.0, "D": 2.0}
    arr = fe_mesh.block_arrays(nodes, tets)
    nid = np.arange(nodes.shape[0])
    ext = nodes.max(axis=0) - nodes.min(axis=0)
    on = lambda ax, v: nid[np.abs(nodes[:, ax] - v) <= 1e-7 * ext[ax]]      # coordinate side sets
    pin = np.concatenate([on(0, nodes[:, 0].min()) * 3, on(1, nodes[:, 1].min()) * 3 + 1,
                          on(2, nodes[:, 2].min()) * 3 + 2])
    load = on(1, nodes[:, 1].max()) * 3 + 1
    bcs = drv.DirichletBCs(np.concatenate([pin, load]),
                           lambda t: np.concatenate([np.zeros(len(pin)), np.full(len(load), 0.01 * t)]))
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy())
    local = dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12, ls_max_evals=100)
    glob = {"max iters": 15, "abs tol": 1e-8, "rel tol": 1e-8}
    return values, nodes, arr, bcs, drv.SparsePattern(ur, uc, arr.n_dofs), scatter, local, glob


def _notch_oracle_run(steps):
    values, nodes, arr, bcs, pattern, scatter, local, glob = _notch_problem()
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **local)
    eq = arr.elem_eq.numpy(); geo = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    rec = {}

    def asm(U, xi_prev):
        o = fe_oracle.assemble_block(prob, eq, U, xi_prev, *geo)
        rec.update(o)
        return o["R"], fe_oracle.coo_dedup_sum(o["K_elem"].reshape(-1), scatter, len(pattern.rows)), o["xi"]
    out = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs), np.zeros((arr.n_elems, 1, 7)),
                                   np.arange(steps + 1.0), glob)
    return out, rec, (values, arr, bcs, pattern, scatter, local, glob)


def test_notch_hosford_deck_converges_over_the_oracle():
    (U_steps, xi, J, logs), rec, _ = _notch_oracle_run(2)
    assert all(l.residual_norms[-1] < 1e-8 * max(l.residual_norms[0], 1.0) or l.residual_norms[-1] < 1e-8 for l in logs)
    assert all(l.iters <= 15 for l in logs)
    alpha = xi[:, 0, 6]
    assert (alpha > 0).mean() > 0.5 and alpha.max() > 3.0 * alpha.mean()     # strain concentrates at the notch
    assert any(a < 1.0 for l in logs for a in l.alphas)                      # the global line search engaged
    assert np.isfinite(U_steps).all()


@pytest.mark.gpu
def test_notch_hosford_deck_cuda_matches_oracle(cuda_device):
    """The reference's example deck at its native size through the CUDA assembler vs the
    oracle assembler: same global Newton history, displacements to 1e-8 (a=100 is
    ill-conditioned: local iteration paths may differ at rounding level, the converged
    states may not)."""
    import torch
    from cmad_b200 import fe, material_from_values
    (Uo, xio, _, lo), rec, (values, arr, bcs, pattern, scatter, local, glob) = _notch_oracle_run(4)
    arr_d = arr.to(cuda_device)
    r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    nw = fe.fe_newton_settings(max_iters=500, abs_tol=1e-12, rel_tol=1e-12, ls_max_evals=100)
    asm = drv.cuda_assembler(material_from_values(values), nw, arr_d, r_plan, k_plan)
    xi0 = torch.zeros((arr.n_elems, 1, 7), dtype=torch.float64, device=cuda_device)
    Ug, xig, _, lg = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs), xi0, np.arange(5.0), glob)
    assert [l.iters for l in lg] == [l.iters for l in lo]
    assert np.abs(Ug - Uo).max() < 1e-8 * np.abs(Uo).max()
    assert np.abs(xig.cpu().numpy() - xio).max() < 1e-8 * np.abs(xio).max()


# ---------------------------------------------------------------- mixed u-p formulation (KA6)
def mixed_uniaxial_cube(div, family="hex8"):
    """tests/fem/test_mixed_up_plastic.py:36-78: unit cube, u-p, symmetry planes on the
    three min faces, the x = max face pulled to ``t``; pressure dofs free."""
    nodes, conn = fe_mesh.structured_hex_mesh((div,) * 3)
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    # the reference's deck driver forces volume degree >= 2 for mixed (cli/common.py:379-391):
    # hex8 x 8 (its default) and tet4 x 4
    arr = fe_mesh.block_arrays(nodes, conn, mixed=True, volume_degree=2)
    nid = np.arange(nodes.shape[0])
    on = lambda ax, v: nid[np.isclose(nodes[:, ax], v)]
    pin = np.concatenate([on(0, 0.0) * 3 + 0, on(1, 0.0) * 3 + 1, on(2, 0.0) * 3 + 2])
    ramp = on(0, 1.0) * 3 + 0
    bcs = drv.DirichletBCs(np.concatenate([pin, ramp]),
                           lambda t: np.concatenate([np.zeros(len(pin)), np.full(len(ramp), t)]))
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy())
    return nodes, arr, bcs, drv.SparsePattern(ur, uc, arr.n_dofs), scatter


def oracle_assembler_mixed(values, arr, scatter, n_unique, record=None):
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **LOCAL_NEWTON)

    def assemble(U, xi_prev):
        o = fe_oracle.assemble_block_mixed(prob, arr.elem_eq.numpy(), arr.elem_eq_p.numpy(), U, xi_prev,
                                           arr.grad_N.numpy(), arr.N.numpy(), arr.det.numpy(), arr.quad_w.numpy(),
                                           arr.h.numpy())
        if record is not None:
            record["sigma"] = o["sigma"]
        vals = np.concatenate([o[k].reshape(-1) for k in ("K_uu", "K_up", "K_pu", "K_pp")])
        return o["R"], fe_oracle.coo_dedup_sum(vals, scatter, n_unique), o["xi"]
    return assemble


def _ka6_targets():
    """Analytic uniaxial J2+Voce state at alpha = 0.05 (MAX_ALPHA, num_steps=2)."""
    mask = np.zeros((3, 3)); mask[0, 0] = 1.0
    stress, strain, _ = analytic.plastic_fields(mask, max_alpha=0.05, num_steps=2)
    return float(strain[0, 0, -1]), float(stress[0, 0, -1])


@pytest.mark.parametrize("family", ["hex8", "tet4"])
def test_mixed_up_uniaxial_cube_known_answer(family):
    """KA6 (tests/fem/test_mixed_up_plastic.py:96-135): sigma_xx = analytic (rtol 1e-5),
    lateral stresses < 1e-4 sigma, p = -sigma/3 - the driver over the oracle's mixed block."""
    values, _, _ = analytic.j2_voce_param_tree("J2")
    axial_strain, sigma_axial = _ka6_targets()
    nodes, arr, bcs, pattern, scatter = mixed_uniaxial_cube(2, family)
    rec = {}
    asm = oracle_assembler_mixed(values, arr, scatter, len(pattern.rows), rec)
    ts = axial_strain * np.arange(6) / 5.0
    U_steps, xi, _, logs = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs),
                                                    np.zeros((arr.n_elems, arr.n_ip, 7)), ts, None, None)
    assert all(l.residual_norms[-1] < 1e-7 for l in logs)
    sig = rec["sigma"]
    np.testing.assert_allclose(sig[..., 0], sigma_axial, rtol=1e-5)
    assert np.abs(sig[..., 3]).max() < 1e-4 * sigma_axial and np.abs(sig[..., 5]).max() < 1e-4 * sigma_axial
    p = U_steps[-1][3 * nodes.shape[0]:]
    np.testing.assert_allclose(p, -sigma_axial / 3.0, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("family,div", [("hex8", 4), ("tet4", 3)])
def test_cuda_mixed_driver_matches_oracle_driver_and_known_answer(cuda_device, family, div):
    import torch
    from cmad_b200 import fe, material_from_values
    values, _, _ = analytic.j2_voce_param_tree("J2")
    axial_strain, sigma_axial = _ka6_targets()
    nodes, arr, bcs, pattern, scatter = mixed_uniaxial_cube(div, family)
    ts = axial_strain * np.arange(6) / 5.0
    arr_d = arr.to(cuda_device)
    r_plan = fe.mixed_r_plan(arr, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    mat, nw = material_from_values(values), fe.fe_newton_settings(**LOCAL_NEWTON)
    asm = drv.cuda_assembler_mixed(mat, nw, arr_d, r_plan, k_plan)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Ug, xig, _, lg = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs), xi0, ts, None, None)
    asm_o = oracle_assembler_mixed(values, arr, scatter, len(pattern.rows))
    Uo, xio, _, lo = drv.fe_quasistatic_drive(asm_o, pattern, bcs, np.zeros(arr.n_dofs),
                                              np.zeros((arr.n_elems, arr.n_ip, 7)), ts, None, None)
    assert [l.iters for l in lg] == [l.iters for l in lo]
    assert np.abs(Ug - Uo).max() < 1e-9 * np.abs(Uo).max()
    assert np.abs(xig.cpu().numpy() - xio).max() < 1e-9 * np.abs(xio).max()
    p = Ug[-1][3 * nodes.shape[0]:]
    np.testing.assert_allclose(p, -sigma_axial / 3.0, rtol=1e-5)
    np.testing.assert_allclose(xig.cpu().numpy()[..., 6], 0.05, rtol=1e-5)


# ---------------------------------------------------------------- mixed u-p FE gradient (config 5 style)
def _mixed_gradient_problem(div=2, family="hex8"):
    from cmad_b200 import Parameters
    from tests.helpers import param_tree
    values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    nodes, arr, bcs, pattern, scatter = mixed_uniaxial_cube(div, family)
    rng = np.random.default_rng(4)
    inner = np.all((nodes > 1e-9) & (nodes < 1 - 1e-9), axis=1)
    nodes = nodes.copy(); nodes[inner] += 0.08 / div * rng.uniform(-1, 1, size=(inner.sum(), 3))
    conn = fe_mesh.structured_hex_mesh((div,) * 3)[1]
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr = fe_mesh.block_arrays(nodes, conn, mixed=True, volume_degree=2)
    return values, P, nodes, arr, bcs, pattern, scatter


def oracle_jvp_mixed(values, P, arr):
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    eq, eqp = arr.elem_eq.numpy(), arr.elem_eq_p.numpy()
    geo = (arr.grad_N.numpy(), arr.N.numpy(), arr.det.numpy(), arr.quad_w.numpy(), arr.h.numpy())
    na = len(prob_eval.active_pid)

    def jvp(U, xi_prev, xi_state, c, dxi_prev, dU):
        dp = np.zeros(na); dp[c] = 1.0
        o = fe_oracle.block_jvp_mixed(prob_eval, eq, eqp, U, xi_prev, xi_state, *geo, dp, dxi_prev, dU=dU)
        dR = np.zeros(arr.n_dofs)
        np.add.at(dR, eq.reshape(-1), o["R_elem"].reshape(-1)); np.add.at(dR, eqp.reshape(-1), o["R_p_elem"].reshape(-1))
        return dR, o["xi"]
    return jvp


def test_mixed_direct_gradient_vs_central_fd_of_the_trajectory():
    """`cmad gradient` on a mixed_plastic.yaml-style deck, restated: dJ/dp through the load
    steps of the mixed u-p formulation by forward sensitivities (K6 over both residual
    blocks - the elastic parameters also enter the pressure rows), against central
    differences of J over re-solved trajectories (all over the oracle assembler)."""
    import copy
    values, P, nodes, arr, bcs, pattern, scatter = _mixed_gradient_problem()
    ts = np.array([0.0, 0.004, 0.008])
    q, dq = _qois(arr, ts)
    tight = {"abs tol": 1e-13, "rel tol": 1e-13, "max iters": 15}
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    J, g = drv.fe_direct_gradient(oracle_assembler_mixed(values, arr, scatter, len(pattern.rows)),
                                  oracle_jvp_mixed(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                  lambda c: z(), ts, 5, tight, q, dq)

    def J_of(v):
        return drv.fe_quasistatic_drive(oracle_assembler_mixed(v, arr, scatter, len(pattern.rows)), pattern, bcs,
                                        np.zeros(arr.n_dofs), z(), ts, tight, q)[2]
    assert abs(J_of(values) - J) < 1e-14 * abs(J)
    paths = [("elastic", "E"), ("elastic", "nu"), ("plastic", "flow stress", "hardening", "voce", "D"),
             ("plastic", "flow stress", "hardening", "voce", "S"), ("plastic", "flow stress", "initial yield", "Y")]
    assert np.abs(g).min() > 0
    for c, path in enumerate(paths):
        def bump(h):
            v = copy.deepcopy(values); d = v
            for k in path[:-1]:
                d = d[k]
            d[path[-1]] = d[path[-1]] * (1 + h)
            return J_of(v)
        h = 1e-5
        d = values
        for k in path:
            d = d[k]
        fd = (bump(h) - bump(-h)) / (2 * h * d)
        assert abs(fd - g[c]) < 5e-5 * abs(g[c]) + 1e-16, (path, fd, g[c])


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["hex8", "tet4"])
def test_cuda_mixed_direct_gradient_matches_oracle(cuda_device, family):
    import torch
    from cmad_b200 import active_param_ids, fe, material_from_values
    values, P, nodes, arr, bcs, pattern, scatter = _mixed_gradient_problem(3, family)
    ts = np.array([0.0, 0.004, 0.008])
    q, dq = _qois(arr, ts)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    Jo, go = drv.fe_direct_gradient(oracle_assembler_mixed(values, arr, scatter, len(pattern.rows)),
                                    oracle_jvp_mixed(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                    lambda c: z(), ts, 5, None, q, dq)
    arr_d = arr.to(cuda_device)
    mat = material_from_values(values)
    r_plan = fe.mixed_r_plan(arr, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    zd = lambda: torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Jg, gg = drv.fe_direct_gradient(
        drv.cuda_assembler_mixed(mat, fe.fe_newton_settings(**LOCAL_NEWTON), arr_d, r_plan, k_plan),
        drv.cuda_jvp_mixed(mat, arr_d, r_plan, active_param_ids(P)), pattern, bcs, np.zeros(arr.n_dofs), zd(),
        lambda c: zd(), ts, 5, None, q, dq)
    assert abs(Jg - Jo) < 1e-10 * abs(Jo)
    assert np.abs(gg - go).max() < 1e-8 * np.abs(go).max(), (gg, go)


# ---------------------------------------------------------------- FE gradient by the discrete adjoint
def oracle_vjp_pair(values, P, arr):
    """``(vjp, vjp_disp)`` of fe_adjoint_gradient over the ORACLE: the JVP oracle is linear in
    (dp, dxi_prev, dU); its transposes are obtained by probing it with unit directions."""
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    eq = arr.elem_eq.numpy(); geo = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    na = len(prob_eval.active_pid)
    shape = (arr.n_elems, arr.n_ip, 7)
    nx = int(np.prod(shape))

    def out_vec(o):
        dR = np.zeros(arr.n_dofs); np.add.at(dR, eq.reshape(-1), o["R_elem"].reshape(-1))
        return dR, o["xi"].reshape(-1)

    def vjp(U, xi_prev, xi_state, Rbar, xibar):
        xb = np.zeros(nx) if xibar is None else np.asarray(xibar).reshape(-1)
        pbar = np.zeros(na); xbp = np.zeros(nx)
        for c in range(na):
            dp = np.zeros(na); dp[c] = 1.0
            dR, dx = out_vec(fe_oracle.block_jvp(prob_eval, eq, U, xi_prev, xi_state, *geo, dp, None))
            pbar[c] = Rbar @ dR + xb @ dx
        for j in range(nx):
            d = np.zeros(nx); d[j] = 1.0
            dR, dx = out_vec(fe_oracle.block_jvp(prob_eval, eq, U, xi_prev, xi_state, *geo, np.zeros(na), d.reshape(shape)))
            xbp[j] = Rbar @ dR + xb @ dx
        return pbar, xbp.reshape(shape)

    def vjp_disp(U, xi_prev, xi_state, xibar):
        xb = np.asarray(xibar).reshape(-1)
        ub = np.zeros(arr.n_dofs)
        for j in range(arr.n_dofs):
            d = np.zeros(arr.n_dofs); d[j] = 1.0
            o = fe_oracle.block_jvp(prob_eval, eq, U, xi_prev, xi_state, *geo, np.zeros(na), None, dU=d)
            ub[j] = xb @ o["xi"].reshape(-1)
        return ub
    return vjp, vjp_disp


def test_adjoint_gradient_equals_direct_gradient_oracle():
    """The discrete adjoint through the load steps (one sparse solve per step) reproduces the
    direct-sensitivity gradient (one per parameter) - both over the oracle."""
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, "tet4")
    ts = np.array([0.0, 0.4, 0.7, 1.0])
    q, dq = _qois(arr, ts)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    asm = oracle_assembler(values, arr, scatter, len(pattern.rows))
    Jd, gd = drv.fe_direct_gradient(asm, oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                    lambda c: z(), ts, 5, None, q, dq)
    vjp, vjp_disp = oracle_vjp_pair(values, P, arr)
    Ja, ga = drv.fe_adjoint_gradient(asm, vjp, vjp_disp, pattern, bcs, np.zeros(arr.n_dofs), z(), ts, 5, None, q, dq)
    assert Ja == Jd
    assert np.abs(ga - gd).max() < 1e-9 * np.abs(gd).max(), (ga, gd)


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["hex8", "tet4"])
@pytest.mark.parametrize("mixed", [False, True])
def test_cuda_adjoint_gradient_matches_direct_and_oracle(cuda_device, family, mixed):
    """`cmad gradient` on an FE deck by the discrete adjoint over the K6 reverse-mode kernels
    (cmadx_fe_block_vjp + cmadx_fe_block_vjp_disp), displacement and mixed u-p: equal to the
    CUDA direct-sensitivity gradient and to the oracle's."""
    import torch
    from cmad_b200 import active_param_ids, fe, material_from_values
    if mixed:
        values, P, nodes, arr, bcs, pattern, scatter = _mixed_gradient_problem(3, family)
        ts = np.array([0.0, 0.004, 0.008])
    else:
        values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(3, family)
        ts = np.array([0.0, 0.4, 0.7, 1.0])
    q, dq = _qois(arr, ts)
    arr_d = arr.to(cuda_device)
    mat = material_from_values(values)
    pid = active_param_ids(P)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    nw = fe.fe_newton_settings(**LOCAL_NEWTON)
    zd = lambda: torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    if mixed:
        r_plan = fe.mixed_r_plan(arr, device=cuda_device)
        asm = drv.cuda_assembler_mixed(mat, nw, arr_d, r_plan, k_plan)
        jvp = drv.cuda_jvp_mixed(mat, arr_d, r_plan, pid)
        vjp, vjp_disp = drv.cuda_vjp(mat, arr_d, pid, stab_mult=1.0)
    else:
        r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
        asm = drv.cuda_assembler(mat, nw, arr_d, r_plan, k_plan)
        jvp = drv.cuda_jvp(mat, arr_d, r_plan, pid)
        vjp, vjp_disp = drv.cuda_vjp(mat, arr_d, pid)
    Jd, gd = drv.fe_direct_gradient(asm, jvp, pattern, bcs, np.zeros(arr.n_dofs), zd(), lambda c: zd(), ts, 5,
                                    None, q, dq)
    Ja, ga = drv.fe_adjoint_gradient(asm, vjp, vjp_disp, pattern, bcs, np.zeros(arr.n_dofs), zd(), ts, 5, None, q, dq)
    assert Ja == Jd
    assert np.abs(ga - gd).max() < 1e-9 * np.abs(gd).max(), (ga, gd)
    # and against the oracle-driven direct gradient
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    if mixed:
        Jo, go = drv.fe_direct_gradient(oracle_assembler_mixed(values, arr, scatter, len(pattern.rows)),
                                        oracle_jvp_mixed(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                        lambda c: z(), ts, 5, None, q, dq)
    else:
        Jo, go = drv.fe_direct_gradient(oracle_assembler(values, arr, scatter, len(pattern.rows)),
                                        oracle_jvp(values, P, arr), pattern, bcs, np.zeros(arr.n_dofs), z(),
                                        lambda c: z(), ts, 5, None, q, dq)
    assert abs(Ja - Jo) < 1e-10 * abs(Jo)
    assert np.abs(ga - go).max() < 1e-8 * np.abs(go).max(), (ga, go)


# ---------------------------------------------------------------- multi-rank FE adjoint gradient (gloo)
def _partitioned(values, P, arr, scatter, n_unique, rank, world):
    """This rank's assemble / vjp / vjp_disp over ITS element range (oracle-backed), with the
    exchanges of the multi-GPU path: all-reduce of R and of the deduplicated K data after the
    assembly, of pbar after the VJP, of the displacement cotangent after vjp_disp; the local
    state and its cotangents stay element-owned (never exchanged)."""
    import torch
    import torch.distributed as dist
    from cmad_b200.objectives import shard_range
    lo, hi = shard_range(arr.n_elems, rank, world)
    sub = arr.slice(lo, hi)
    n2 = (3 * arr.n_basis) ** 2
    asm_local = oracle_assembler(values, sub, scatter[lo * n2:hi * n2], n_unique)
    vjp_l, vjp_disp_l = oracle_vjp_pair(values, P, sub)

    def allsum(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    def assemble(U, xi_prev):
        R, K, xi = asm_local(U, xi_prev)
        return allsum(R), allsum(K), xi

    def vjp(U, xi_prev, xi_state, Rbar, xibar):
        pbar, xbp = vjp_l(U, xi_prev, xi_state, Rbar, xibar)
        return allsum(pbar), xbp

    def vjp_disp(U, xi_prev, xi_state, xibar):
        return allsum(vjp_disp_l(U, xi_prev, xi_state, xibar))
    return assemble, vjp, vjp_disp, (lo, hi)


def _fe_adjoint_gloo_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, "tet4")
    ts = np.array([0.0, 0.5, 1.0])
    q, dq = _qois(arr, ts)
    assemble, vjp, vjp_disp, (lo, hi) = _partitioned(values, P, arr, scatter, len(pattern.rows), rank, world)
    z = np.zeros((hi - lo, arr.n_ip, 7))
    J, g = drv.fe_adjoint_gradient(assemble, vjp, vjp_disp, pattern, bcs, np.zeros(arr.n_dofs), z, ts, 5, None, q, dq)
    ret[rank] = (J, g, lo, hi)
    dist.destroy_process_group()


def test_two_rank_gloo_fe_adjoint_gradient_equals_single_process():
    """Element partition over 2 ranks: every rank runs the same global Newton / adjoint sweep on
    all-reduced (R, K, pbar, ubar) while its local states and their cotangents stay rank-local;
    both ranks obtain the single-process gradient."""
    import os
    import torch.multiprocessing as tmp
    values, P, nodes, arr, bcs, pattern, scatter = _gradient_problem(2, "tet4")
    ts = np.array([0.0, 0.5, 1.0])
    q, dq = _qois(arr, ts)
    z = lambda: np.zeros((arr.n_elems, arr.n_ip, 7))
    vjp, vjp_disp = oracle_vjp_pair(values, P, arr)
    J1, g1 = drv.fe_adjoint_gradient(oracle_assembler(values, arr, scatter, len(pattern.rows)), vjp, vjp_disp,
                                     pattern, bcs, np.zeros(arr.n_dofs), z(), ts, 5, None, q, dq)
    ret = tmp.Manager().dict()
    tmp.spawn(_fe_adjoint_gloo_worker, args=(2, 29100 + os.getpid() % 1500, ret), nprocs=2, join=True)
    assert ret[0][2:] == (0, arr.n_elems // 2) and ret[1][3] == arr.n_elems
    for r in (0, 1):
        assert abs(ret[r][0] - J1) < 1e-12 * abs(J1)
        assert np.abs(ret[r][1] - g1).max() < 1e-9 * np.abs(g1).max(), (ret[r][1], g1)
