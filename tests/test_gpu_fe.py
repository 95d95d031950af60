"""GPU parity tests of the FE element-block kernels (K3/K4) and the deterministic
segment sums (K5), through the C-ABI, against the CPU block oracle on the same
seeded inputs: values within 1e-10 (relative to the largest entry of each
quantity), Newton iteration counts and branch flags exactly equal."""
import numpy as np
import pytest
import torch

from cmad_b200 import fe, fe_mesh, material_from_values
from oracle import analytic, fe_oracle, oracle_c as oc
from tests.helpers import param_tree, rel_err, rotation_matrix

pytestmark = pytest.mark.gpu
TOL = 1e-10
NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)


def _mesh(family, divisions, distort=0.05, seed=1):
    nodes, conn = fe_mesh.structured_hex_mesh(divisions)
    rng = np.random.default_rng(seed)
    h = 1.0 / max(divisions)
    nodes = nodes + distort * h * rng.uniform(-1, 1, size=nodes.shape)
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    return nodes, conn


def _check_block(cuda_device, values, family, divisions, steps=3, force_generic=False, newton=None,
                 ramp=0.004, exact=True):
    newton = newton or NEWTON
    nodes, conn = _mesh(family, divisions)
    arr_h = fe_mesh.block_arrays(nodes, conn)
    arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(force_generic=force_generic, **newton)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **newton)
    n_e, n_ip = arr.n_elems, arr.n_ip
    xi_ref = np.zeros((n_e, n_ip, 7))
    xi = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    plastic = False
    for s in range(1, steps + 1):
        U = fe_mesh.synthetic_displacement(nodes, t=float(s), seed=40 + s, ramp=ramp, noise=4e-4)
        out = fe.fe_block_launch(mat, nw, arr, torch.from_numpy(U).to(cuda_device), xi,
                                 outputs=("xi", "R_elem", "K_elem", "sigma", "iters", "flags"))
        ref = fe_oracle.assemble_block(prob, arr_h.elem_eq.numpy(), U, xi_ref, arr_h.grad_N.numpy(),
                                       arr_h.det.numpy(), arr_h.quad_w.numpy())
        torch.cuda.synchronize()
        it, fl = out["iters"].cpu().numpy(), out["flags"].cpu().numpy()
        if exact:
            assert np.array_equal(it, ref["iters"]), np.argwhere(it != ref["iters"])[:5]
            assert np.array_equal(fl, ref["flags"]), np.argwhere(fl != ref["flags"])[:5]
        else:
            assert ((it == ref["iters"]) & (fl == ref["flags"])).mean() > 0.999
        for k in ("xi", "sigma", "R_elem", "K_elem"):
            assert rel_err(out[k].cpu().numpy(), ref[k]) < TOL, (k, s, rel_err(out[k].cpu().numpy(), ref[k]))
        xi, xi_ref = out["xi"], ref["xi"]
        plastic |= bool((ref["flags"] & 2).any())
    assert plastic
    return out, ref, arr, arr_h


@pytest.mark.parametrize("family,divisions", [("tet4", (5, 4, 3)), ("hex8", (6, 5, 5))])
@pytest.mark.parametrize("kind", ["J2", "J2-generic", "hill", "hosford"])
def test_block_parity_vs_oracle(cuda_device, family, divisions, kind):
    """Ragged element counts (not multiples of the block size), three load steps
    with the state carried, every yield surface; J2 through the radial-return
    kernel and through the generic 7x7 Newton."""
    if kind.startswith("J2"):
        values, _, _ = param_tree("J2")
    elif kind == "hill":
        values, _, _ = param_tree("hill", hill=(0.45, 0.55, 0.5, 1.4, 1.5, 1.6))
    else:
        values, _, _ = param_tree("hosford", a=6.0)
    _check_block(cuda_device, values, family, divisions, force_generic=(kind == "J2-generic"))


@pytest.mark.parametrize("family", ["tet4", "hex8"])
def test_block_parity_rotated_axes_and_linear_hardening(cuda_device, family):
    Q = rotation_matrix([1.0, 2.0, -0.5], 0.7)
    values, _, _ = param_tree("hill", hill=(0.45, 0.55, 0.5, 1.4, 1.5, 1.6),
                              hardening=("voce", "linear"), rotation=Q)
    _check_block(cuda_device, values, family, (3, 3, 2))


@pytest.mark.parametrize("family", ["tet4", "hex8"])
def test_residual_only_scatter_and_dedup(cuda_device, family):
    """K4 reproduces K3's residual; the deterministic segment sums equal
    the oracle's sequential scatter-add (R) and COO dedup (K); the atomic R path
    agrees to rounding; two runs of the deterministic path are bit-identical."""
    values, _, _ = param_tree("J2")
    out, ref, arr, arr_h = _check_block(cuda_device, values, family, (4, 3, 3), steps=2)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    U = torch.from_numpy(fe_mesh.synthetic_displacement(
        fe_mesh.structured_hex_mesh((4, 3, 3))[0], 2.0, seed=42, ramp=0.004, noise=4e-4)).to(cuda_device)
    xi_prev = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    eq = arr_h.elem_eq.numpy()
    r_plan = fe.SegmentPlan(eq.reshape(-1), arr.n_dofs, device=cuda_device)
    ur, uc, scatter = fe_mesh.coo_dedup(eq)
    k_plan = fe.SegmentPlan(scatter, len(ur), device=cuda_device)

    R1, vals, xi1 = fe.assemble_element_block(mat, nw, arr, U, xi_prev, r_plan=r_plan)
    R2, vals2, _ = fe.assemble_element_block(mat, nw, arr, U, xi_prev, r_plan=r_plan)
    Ra, _, _ = fe.assemble_element_block(mat, nw, arr, U, xi_prev)            # atomics
    Rk4 = fe.assemble_element_block_residual(mat, nw, arr, U, xi_prev, r_plan=r_plan)
    Kd = k_plan.sum(vals)
    torch.cuda.synchronize()
    assert torch.equal(R1, R2) and torch.equal(vals, vals2)
    assert rel_err(Rk4.cpu().numpy(), R1.cpu().numpy()) < 1e-13
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    o = fe_oracle.assemble_block(prob, eq, U.cpu().numpy(), xi_prev.cpu().numpy(), arr_h.grad_N.numpy(),
                                 arr_h.det.numpy(), arr_h.quad_w.numpy())
    assert rel_err(R1.cpu().numpy(), o["R"]) < TOL
    assert rel_err(Ra.cpu().numpy(), o["R"]) < TOL
    Kref = fe_oracle.coo_dedup_sum(o["K_elem"].reshape(-1), scatter, len(ur))
    assert rel_err(Kd.cpu().numpy(), Kref) < TOL
    # accumulate mode adds a second block's contribution into an existing R
    ones = torch.ones(r_plan.n_items, dtype=torch.float64, device=cuda_device)
    R3 = r_plan.sum(ones, out=R1.clone(), accumulate=True)
    cnt = np.bincount(eq.reshape(-1), minlength=arr.n_dofs)
    assert np.allclose(R3.cpu().numpy(), R1.cpu().numpy() + cnt)
    r_plan.close(); k_plan.close()


def test_known_answers_ka2_tet_and_ka3_hex(cuda_device):
    """KA2 (tests/global_residuals/test_for_model_coupled.py:34-82, 231-295): tet
    barycentre fixture, plastic, alpha > 0.  KA3 (tests/fem/test_per_element_coupled.py:
    80-91, 174-386): reference hex, plastic loading.  The CUDA kernels reproduce the
    oracle (which the CPU suite pins on the reference's assertions for both)."""
    values, _, _ = analytic.j2_voce_param_tree("J2")
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    # KA2
    nodes = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    arr_h = fe_mesh.block_arrays(nodes, np.array([[0, 1, 2, 3]]))
    assert np.allclose(arr_h.grad_N[0, 0].numpy(), [[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    assert np.isclose(float(arr_h.det[0, 0] * arr_h.quad_w[0]), 1.0 / 6.0)
    U = np.zeros((4, 3)); U[1, 0] = .005; U[2, 1] = .003; U[3, 2] = .002
    cases = [(arr_h, U.reshape(-1))]
    # KA3
    hn, hc = fe_mesh.structured_hex_mesh((1, 1, 1), (2.0, 2.0, 2.0), origin=(-1.0, -1.0, -1.0))
    Uh = np.zeros((8, 3))
    Uh[[1, 2, 5, 6], 0] = 0.005; Uh[[2, 3, 6, 7], 1] = 0.003; Uh[[4, 5, 6, 7], 2] = 0.002
    Ug = np.zeros((8, 3)); Ug[hc[0]] = Uh
    cases.append((fe_mesh.block_arrays(hn, hc), Ug.reshape(-1)))
    for a_h, Uc in cases:
        a = a_h.to(cuda_device)
        xi0 = torch.zeros((1, a.n_ip, 7), dtype=torch.float64, device=cuda_device)
        out = fe.fe_block_launch(mat, nw, a, torch.from_numpy(Uc).to(cuda_device), xi0,
                                 outputs=("xi", "R_elem", "K_elem", "iters", "flags"))
        ref = fe_oracle.assemble_block(prob, a_h.elem_eq.numpy(), Uc, np.zeros((1, a.n_ip, 7)),
                                       a_h.grad_N.numpy(), a_h.det.numpy(), a_h.quad_w.numpy())
        torch.cuda.synchronize()
        assert (out["xi"][0, :, 6] > 0).any() and (out["flags"] & 2).any()
        for k in ("xi", "R_elem", "K_elem"):
            assert rel_err(out[k].cpu().numpy(), ref[k]) < TOL, k
        assert np.array_equal(out["iters"].cpu().numpy(), ref["iters"])
        # K_e symmetric (associative plasticity) and R_e self-equilibrated
        K = out["K_elem"][0].cpu().numpy()
        assert np.abs(K - K.T).max() < 1e-9 * np.abs(K).max()
        assert np.abs(out["R_elem"][0].cpu().numpy().reshape(-1, 3).sum(axis=0)).max() < 1e-9


@pytest.mark.parametrize("family", ["tet4", "hex8"])
def test_zero_displacement_is_elastic_and_finite(cuda_device, family):
    """U = 0 (first global Newton iterate of a run): zero deviator, where the J2
    normal is NaN and the reference's jnp.where masks it (paths.py:27): R = 0,
    xi = xi_prev, and K_e is the finite elastic stiffness."""
    values, _, _ = param_tree("J2")
    nodes, conn = _mesh(family, (2, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    U = np.zeros(arr.n_dofs)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    for generic in (False, True):
        nw = fe.fe_newton_settings(force_generic=generic, **NEWTON)
        out = fe.fe_block_launch(mat, nw, arr, torch.from_numpy(U).to(cuda_device), xi0,
                                 outputs=("xi", "R_elem", "K_elem", "iters", "flags"))
        ref = fe_oracle.assemble_block(prob, arr_h.elem_eq.numpy(), U, xi0.cpu().numpy(), arr_h.grad_N.numpy(),
                                       arr_h.det.numpy(), arr_h.quad_w.numpy())
        torch.cuda.synchronize()
        assert torch.isfinite(out["K_elem"]).all() and torch.isfinite(out["xi"]).all()
        assert float(out["R_elem"].abs().max()) == 0.0 and float(out["xi"].abs().max()) == 0.0
        assert int(out["iters"].max()) == 0 and int(out["flags"].max()) == 0
        assert rel_err(out["K_elem"].cpu().numpy(), ref["K_elem"]) < TOL


def test_empty_block_and_argument_errors(cuda_device):
    values, _, _ = param_tree("J2")
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    nodes, conn = _mesh("tet4", (1, 1, 1))
    arr = fe_mesh.block_arrays(nodes, conn).to(cuda_device)
    U = torch.zeros(arr.n_dofs, dtype=torch.float64, device=cuda_device)
    empty = arr.slice(0, 0)
    out = fe.fe_block_launch(mat, nw, empty, U, torch.zeros((0, 1, 7), dtype=torch.float64, device=cuda_device))
    assert out["K_elem"].shape == (0, 12, 12)
    with pytest.raises(ValueError):
        fe.fe_block_launch(mat, nw, arr, U[:-1].contiguous(), torch.zeros((6, 1, 7), dtype=torch.float64, device=cuda_device))
    with pytest.raises(ValueError):
        fe.fe_block_launch(mat, nw, arr, U, torch.zeros((6, 2, 7), dtype=torch.float64, device=cuda_device))
    # a rule with more points than any the reference tabulates (> 64) is rejected, not mis-assembled
    bad = fe_mesh.FEBlockArrays(arr.elem_eq, arr.grad_N.repeat(1, 65, 1, 1).contiguous(), arr.det.repeat(1, 65).contiguous(),
                                arr.quad_w.repeat(65).contiguous(), arr.N, arr.n_dofs)
    with pytest.raises(ValueError):
        fe.fe_block_launch(mat, nw, bad, U, torch.zeros((6, 65, 7), dtype=torch.float64, device=cuda_device))
    # Elastic model blocks are CLOSED_FORM in the reference (cli/common.py:347-354): not this kernel
    with pytest.raises(NotImplementedError):
        fe.fe_block_launch(material_from_values({"elastic": {"kappa": 100.0, "mu": 50.0}}, model="elastic"),
                           nw, arr, U, torch.zeros((6, 1, 7), dtype=torch.float64, device=cuda_device))


@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
def test_block_jvp_vs_oracle(cuda_device, family, kind):
    """K6: forward sensitivities of (xi, R_e) w.r.t. (params, xi_prev) at fixed U, at
    the converged state of a primal K4 launch, against the IFT oracle (itself checked
    against central differences of the Newton-running block in the CPU suite)."""
    from cmad_b200 import Parameters, active_param_ids
    if kind == "J2":
        values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    elif kind == "hill":
        values, act, tr = param_tree("hill", hill=(0.45, 0.55, 0.5, 1.4, 1.5, 1.6),
                                     active=("E", "nu", "D", "S", "Y") + tuple("FGHLMN"))
    else:
        values, act, tr = param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    pid = active_param_ids(P)
    nodes, conn = _mesh(family, (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    rng = np.random.default_rng(2)
    n_e, n_ip = arr.n_elems, arr.n_ip
    U1 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 1.0, seed=1, ramp=0.004, noise=4e-4)).to(cuda_device)
    U2 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 2.0, seed=2, ramp=0.004, noise=4e-4)).to(cuda_device)
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    xi1 = fe.fe_block_launch(mat, nw, arr, U1, xi0, ("xi",))["xi"]
    prim = fe.fe_block_launch(mat, nw, arr, U2, xi1, ("xi", "R_elem", "flags"))
    assert bool((prim["flags"] & 2).any()) and not bool((prim["flags"] & 2).all() and False)
    dp = rng.standard_normal(len(pid)) * np.array([3e3, 0.01, 1.5, 7.0, 4.0] + [0.05] * (len(pid) - 5))
    dxp = torch.from_numpy(1e-4 * rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    out = fe.fe_block_jvp(mat, arr, U2, xi1, prim["xi"], pid, dp, dxp, outputs=("xi", "R_elem", "R_global"))
    out0 = fe.fe_block_jvp(mat, arr, U2, xi1, prim["xi"], pid, dp, None)
    torch.cuda.synchronize()
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    geo = (arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy())
    eq = arr_h.elem_eq.numpy()
    ref = fe_oracle.block_jvp(prob_eval, eq, U2.cpu().numpy(), xi1.cpu().numpy(), prim["xi"].cpu().numpy(),
                              *geo, dp, dxp.cpu().numpy())
    ref0 = fe_oracle.block_jvp(prob_eval, eq, U2.cpu().numpy(), xi1.cpu().numpy(), prim["xi"].cpu().numpy(),
                               *geo, dp, None)
    # with a displacement direction as well (total state sensitivity, dR includes K dU)
    dUv = torch.from_numpy(1e-4 * rng.standard_normal(arr.n_dofs)).to(cuda_device)
    outU = fe.fe_block_jvp(mat, arr, U2, xi1, prim["xi"], pid, dp, dxp, dU=dUv)
    refU = fe_oracle.block_jvp(prob_eval, eq, U2.cpu().numpy(), xi1.cpu().numpy(), prim["xi"].cpu().numpy(),
                               *geo, dp, dxp.cpu().numpy(), dU=dUv.cpu().numpy())
    Kd = fe.fe_block_launch(mat, nw, arr, U2, xi1, ("xi", "K_elem"))["K_elem"]
    KdU = torch.einsum("eij,ej->ei", Kd, dUv[arr.elem_eq.long()])
    assert rel_err((outU["R_elem"] - out["R_elem"]).cpu().numpy(), KdU.cpu().numpy()) < 1e-8
    for o, r in ((out, ref), (out0, ref0), (outU, refU)):
        assert rel_err(o["xi"].cpu().numpy(), r["xi"]) < 1e-9
        assert rel_err(o["R_elem"].cpu().numpy(), r["R_elem"]) < 1e-9
    Rg = np.zeros(arr.n_dofs); np.add.at(Rg, eq.reshape(-1), ref["R_elem"].reshape(-1))
    assert rel_err(out["R_global"].cpu().numpy(), Rg) < 1e-9
    # linearity in the direction (a size-independent property of the JVP)
    out2 = fe.fe_block_jvp(mat, arr, U2, xi1, prim["xi"], pid, 2.0 * dp, 2.0 * dxp)
    assert rel_err(out2["R_elem"].cpu().numpy(), 2.0 * out["R_elem"].cpu().numpy()) < 1e-12
    with pytest.raises(NotImplementedError):
        fe.fe_block_jvp(mat, arr, U2, xi1, prim["xi"], [_lib_pid_q00()], [1.0])


def _lib_pid_q00():
    from cmad_b200 import _lib
    return _lib.P_Q00


@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("kind", ["J2", "hill-rotated", "hosford"])
def test_block_vjp_is_the_transpose_of_the_jvp(cuda_device, family, kind):
    """K6 reverse mode against the ORACLE's JVP through the adjoint identity
    <Rbar, dR> + <xibar, dxi> == <pbar, dp> + <xibar_prev, dxi_prev> for random directions
    and cotangents, plus bit-reproducibility of the reduced parameter gradient."""
    from cmad_b200 import Parameters, active_param_ids
    if kind == "J2":
        values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    elif kind == "hosford":
        values, act, tr = param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y"))
    else:
        values, act, tr = param_tree("hill", hill=(0.45, 0.55, 0.5, 1.4, 1.5, 1.6),
                                     active=("E", "nu", "D", "S", "Y") + tuple("FGHLMN"),
                                     rotation=rotation_matrix([1.0, 2.0, -0.5], 0.7))
    P = Parameters(values, act, tr)
    pid = active_param_ids(P)
    nodes, conn = _mesh(family, (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    rng = np.random.default_rng(7)
    n_e, n_ip = arr.n_elems, arr.n_ip
    U1 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 1.0, seed=1, ramp=0.004, noise=4e-4)).to(cuda_device)
    U2 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 2.0, seed=2, ramp=0.004, noise=4e-4)).to(cuda_device)
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    xi1 = fe.fe_block_launch(mat, nw, arr, U1, xi0, ("xi",))["xi"]
    prim = fe.fe_block_launch(mat, nw, arr, U2, xi1, ("xi", "flags"))
    assert bool((prim["flags"] & 2).any())
    Rbar = torch.from_numpy(rng.standard_normal(arr.n_dofs)).to(cuda_device)
    xibar = torch.from_numpy(rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    pbar, xbp = fe.fe_block_vjp(mat, arr, U2, xi1, prim["xi"], pid, Rbar, xibar)
    pbar2, xbp2 = fe.fe_block_vjp(mat, arr, U2, xi1, prim["xi"], pid, Rbar, xibar)
    torch.cuda.synchronize()
    assert torch.equal(pbar, pbar2) and torch.equal(xbp, xbp2)
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    geo = (arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy())
    eq = arr_h.elem_eq.numpy()
    scale = np.array([3e3, 0.01, 1.5, 7.0, 4.0] + [0.05] * (len(pid) - 5))
    for trial in range(3):
        dp = rng.standard_normal(len(pid)) * scale
        dxp = 1e-4 * rng.standard_normal((n_e, n_ip, 7))
        jv = fe_oracle.block_jvp(prob_eval, eq, U2.cpu().numpy(), xi1.cpu().numpy(), prim["xi"].cpu().numpy(),
                                 *geo, dp, dxp)
        dR = np.zeros(arr.n_dofs); np.add.at(dR, eq.reshape(-1), jv["R_elem"].reshape(-1))
        lhs = float(Rbar.cpu().numpy() @ dR + (xibar.cpu().numpy() * jv["xi"]).sum())
        rhs = float(pbar.cpu().numpy() @ dp + (xbp.cpu().numpy() * dxp).sum())
        terms = abs(Rbar.cpu().numpy() @ dR) + abs((xibar.cpu().numpy() * jv["xi"]).sum())
        assert abs(lhs - rhs) < 1e-9 * terms, (trial, lhs, rhs)
    # without a state cotangent, and with no active parameters
    pbar0, xbp0 = fe.fe_block_vjp(mat, arr, U2, xi1, prim["xi"], pid, Rbar, None)
    pnone, xbpn = fe.fe_block_vjp(mat, arr, U2, xi1, prim["xi"], [], Rbar, None)
    assert pnone.numel() == 0 and torch.equal(xbp0, xbpn)


@pytest.mark.parametrize("family,div", [("hex8", 24), ("tet4", 16)])
def test_repeated_launches_are_bit_identical(cuda_device, family, div):
    """No atomics on the deterministic path and every shared-memory hand-over inside the
    hex8 kernel is warp-synchronised: five launches over ~14-25 k elements (hundreds of
    resident blocks in flight) give bit-identical K_e, R_e, xi.  (compute-sanitizer is
    closed on this GPU pool, so racecheck is replaced by this and the exact-parity tests.)"""
    values, _, _ = param_tree("J2")
    nodes, conn = _mesh(family, (div, div, div))
    arr = fe_mesh.block_arrays(nodes, conn, device=cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    U = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 2.0, seed=8, ramp=0.004, noise=1e-3 / div)).to(cuda_device)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    ref = None
    for rep in range(5):
        o = fe.fe_block_launch(mat, nw, arr, U, xi0, ("xi", "R_elem", "K_elem", "flags"))
        torch.cuda.synchronize()
        if ref is None:
            ref = {k: v.clone() for k, v in o.items()}
            assert bool((ref["flags"] & 2).any())
        else:
            for k in ref:
                assert torch.equal(o[k], ref[k]), (k, rep)
    # K_e symmetric and R_e self-equilibrated on every element
    K = ref["K_elem"]
    assert float((K - K.transpose(1, 2)).abs().max()) < 1e-9 * float(K.abs().max())
    Rsum = ref["R_elem"].reshape(arr.n_elems, -1, 3).sum(dim=1).abs().max()
    assert float(Rsum) < 1e-9 * float(ref["R_elem"].abs().max())


# ---------------------------------------------------------------- post-processing / QoI / embedded BCs
@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("kind", ["J2", "hill_rot"])
def test_cauchy_at_ips_and_reaction(cuda_device, family, kind):
    """evaluate_cauchy_at_ips (postprocess.py:35-185) and FELoadMatch._reaction_at
    (fe_load_match.py:179-196) on the CUDA path vs the oracle."""
    from cmad_b200 import fe, fe_mesh, material_from_values
    from oracle import fe_oracle
    from tests.golden.materials import material
    values = material(kind)
    nodes, conn = fe_mesh.structured_hex_mesh((3, 2, 2))
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    arr_h = fe_mesh.block_arrays(nodes, conn)
    arr = arr_h.to(cuda_device)
    U = fe_mesh.synthetic_displacement(nodes, 2.0, seed=11, ramp=0.004, noise=5e-4)
    mat, nw = material_from_values(values), fe.fe_newton_settings()
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Ud = torch.from_numpy(U).to(cuda_device)
    o = fe.fe_block_launch(mat, nw, arr, Ud, xi0, ("xi", "sigma", "R_elem"))
    sig = fe.evaluate_cauchy_at_ips(mat, arr, Ud, o["xi"])
    torch.cuda.synchronize()
    # the kernel that solved the step and the post-processing kernel agree...
    assert rel_err(sig.cpu().numpy(), o["sigma"].cpu().numpy()) < 1e-12
    # ...and both agree with the oracle's model.cauchy at the stored state
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, max_iters=0)
    ref = fe_oracle.cauchy_at_ips(prob, arr_h.elem_eq.numpy(), U, o["xi"].cpu().numpy(), arr_h.grad_N.numpy())
    assert rel_err(sig.cpu().numpy(), ref) < 1e-10
    # reactions: residual summed over the x = max face dofs, per component
    nid = np.arange(nodes.shape[0])[np.isclose(nodes[:, 0], nodes[:, 0].max())]
    eqs = [nid * 3 + c for c in range(3)]
    plan = fe.SegmentPlan(arr_h.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    R = plan.sum(o["R_elem"].reshape(-1))
    react = fe.ReactionPlan(eqs, cuda_device)(R).cpu().numpy()
    Rn = R.cpu().numpy()
    assert np.allclose(react, [Rn[e].sum() for e in eqs], rtol=1e-12, atol=1e-12 * np.abs(Rn).max())


def test_device_embedded_bcs_match_reference_restatement(cuda_device):
    """_embedded_bc_enforce + _embedded_residual (sparse_solve.py:1058-1174) on the device vs the
    NumPy restatement and vs the host treatment the driver uses; then the same Newton
    trajectory with device-side enforcement as with host-side enforcement."""
    from cmad_b200 import fe, fe_driver as drv, fe_mesh, material_from_values
    from oracle import analytic, fe_oracle
    from tests.test_fe_driver import LOCAL_NEWTON, uniaxial_cube
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, arr, bcs, pattern, scatter = uniaxial_cube(4, "hex8")
    arr_d = arr.to(cuda_device)
    r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=cuda_device)
    mat, nw = material_from_values(values), fe.fe_newton_settings(**LOCAL_NEWTON)
    asm = drv.cuda_assembler(mat, nw, arr_d, r_plan, k_plan)
    rng = np.random.default_rng(0)
    U = fe_mesh.synthetic_displacement(nodes, 1.5, seed=2, ramp=0.003, noise=3e-4)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    R, K, _ = asm.device(U, xi0)
    t = 0.7
    vals = bcs.values(t)
    plan = fe.EmbeddedBCPlan(pattern.rows, pattern.cols, pattern.n, bcs.indices, device=cuda_device)
    r_d, K_d = plan.apply(K, R, torch.from_numpy(U).to(cuda_device), torch.from_numpy(vals).to(cuda_device))
    r_ref, K_ref = fe_oracle.embedded_system(pattern.rows, pattern.cols, pattern.n, K.cpu().numpy(), R.cpu().numpy(),
                                             U, bcs.indices, vals)
    assert rel_err(r_d.cpu().numpy(), r_ref) < 1e-13 and rel_err(K_d.cpu().numpy(), K_ref) == 0.0
    r_h, K_h = drv.embedded_system(pattern, K.cpu().numpy(), R.cpu().numpy(), U, bcs, t)
    assert rel_err(r_h, r_ref) < 1e-13
    assert abs(K_h - pattern.csr(K_ref).tocsc()).max() < 1e-13 * np.abs(K_ref).max()
    # whole load-step loop
    ts = np.linspace(0.0, 1.0, 4)
    Uh, xih, _, lh = drv.fe_quasistatic_drive(asm, pattern, bcs, np.zeros(arr.n_dofs), xi0, ts)
    asm_dev = drv.DeviceEmbeddedBCs(asm, pattern, bcs, cuda_device)
    Ud, xid, _, ld = drv.fe_quasistatic_drive(asm_dev, pattern, bcs, np.zeros(arr.n_dofs), xi0, ts)
    assert [l.iters for l in ld] == [l.iters for l in lh]
    assert np.abs(Ud - Uh).max() < 1e-12 * np.abs(Uh).max()
    assert np.abs(xid.cpu().numpy() - xih.cpu().numpy()).max() < 1e-12


@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("mixed", [False, True])
def test_closed_form_elastic_blocks(cuda_device, family, mixed):
    """CLOSED_FORM blocks of the `Elastic` model (kappa, mu given: examples/mixed_elastic.yaml;
    KA4's material kappa = 100, mu = 50): R_e and K_e from the element kernels' elastic branch vs
    the closed-form stress kappa tr(eps) I + 2 mu dev(eps) (elastic_stress.py:24-42) assembled
    with NumPy; K_e constant = B^T Cel B; the u-p variant against the mixed oracle."""
    values = {"elastic": {"kappa": 100.0, "mu": 50.0}}
    nodes, conn = _mesh(family, (2, 2, 1), distort=0.08, seed=4)
    arr_h = fe_mesh.block_arrays(nodes, conn, mixed=mixed)
    arr = arr_h.to(cuda_device)
    rng = np.random.default_rng(3)
    U = 0.05 * rng.standard_normal(nodes.shape[0] * 3)
    if mixed:
        U = np.concatenate([U, 3.0 * rng.standard_normal(nodes.shape[0])])
    mat = fe.closed_form_elastic_material(values)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    Ud = torch.from_numpy(U).to(cuda_device)
    eq = arr_h.elem_eq.numpy()
    gN, det, w = arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy()
    n_e, n_ip, n_b, _ = gN.shape
    U_e = U[eq].reshape(n_e, n_b, 3)
    gu = np.einsum("eak,epaj->epkj", U_e, gN)
    eps = 0.5 * (gu + np.swapaxes(gu, 2, 3))
    tr = np.trace(eps, axis1=2, axis2=3)
    I3 = np.eye(3)
    sig = 100.0 * tr[..., None, None] * I3 + 2 * 50.0 * (eps - tr[..., None, None] / 3.0 * I3)
    wdv = det * w[None, :]
    if not mixed:
        o = fe.fe_block_launch(mat, fe.fe_newton_settings(), arr, Ud, xi0, ("xi", "R_elem", "K_elem", "iters"))
        torch.cuda.synchronize()
        R_ref = np.einsum("epaj,epji,ep->eai", gN, sig, wdv).reshape(n_e, -1)
        assert rel_err(o["R_elem"].cpu().numpy(), R_ref) < 1e-12
        assert float(o["xi"].abs().max()) == 0.0 and int(o["iters"].max()) == 0
        lam = 100.0 - 2 * 50.0 / 3.0
        Cel = lam * np.einsum("ji,kl->jikl", I3, I3) + 50.0 * (np.einsum("jk,il->jikl", I3, I3) + np.einsum("jl,ik->jikl", I3, I3))
        K_ref = np.einsum("epaj,jikl,epbl,ep->eaibk", gN, Cel, gN, wdv).reshape(n_e, 3 * n_b, 3 * n_b)
        assert rel_err(o["K_elem"].cpu().numpy(), K_ref) < 1e-12
    else:
        R, vals, xi = fe.assemble_element_block_mixed(mat, fe.fe_newton_settings(), arr, Ud, xi0,
                                                      r_plan=fe.mixed_r_plan(arr))
        torch.cuda.synchronize()
        prob = oc.describe({"rotation matrix": np.eye(3), "elastic": values["elastic"],
                            "plastic": {"effective stress": {"J2": 0.0},
                                        "flow stress": {"initial yield": {"Y": 1e300}, "hardening": {}}}},
                           None, newton_mode="traced", strain_comps=9, **NEWTON)
        ref = fe_oracle.assemble_block_mixed(prob, eq, arr_h.elem_eq_p.numpy(), U, np.zeros((n_e, n_ip, 7)), gN,
                                             arr_h.N.numpy(), det, w, arr_h.h.numpy())
        assert rel_err(R.cpu().numpy(), ref["R"]) < 1e-12
        v_ref = np.concatenate([ref[k].reshape(-1) for k in ("K_uu", "K_up", "K_pu", "K_pp")])
        assert rel_err(vals.cpu().numpy(), v_ref) < 1e-12
        # momentum residual of the closed form: dev(sigma) - p I
        p_ip = np.einsum("pa,ea->ep", arr_h.N.numpy(), U[arr_h.elem_eq_p.numpy()])
        sdev = sig - np.trace(sig, axis1=2, axis2=3)[..., None, None] / 3.0 * I3 - p_ip[..., None, None] * I3
        R_u = np.zeros(U.size); np.add.at(R_u, eq.reshape(-1), np.einsum("epaj,epji,ep->eai", gN, sdev, wdv).reshape(-1))
        n_u = 3 * nodes.shape[0]
        assert rel_err(R.cpu().numpy()[:n_u], R_u[:n_u]) < 1e-12


@pytest.mark.parametrize("family", ["tet4", "hex8"])
def test_fe_two_pass_deferral_keeps_iterates_and_counts(cuda_device, family):
    """K3 with the generic Newton in two passes (elements with a point needing more than K
    updates go to a compacted second launch) vs a single pass: identical xi / counts / flags, R_e
    and K_e equal to rounding (the two launches are separately compiled instantiations)."""
    from tests.golden.materials import material
    values = material("hosford_notch")                                  # near-Tresca: 0 / 2 / 5-10+ updates
    nodes, conn = _mesh(family, (6, 6, 6), distort=0.05, seed=8)
    arr = fe_mesh.block_arrays(nodes, conn).to(cuda_device)
    U = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 1.0, seed=3, ramp=0.004, noise=6e-4)).to(cuda_device)
    mat = material_from_values(values)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    kw = dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12, ls_max_evals=100)
    outs = ("xi", "R_elem", "K_elem", "iters", "flags")
    base = fe.fe_block_launch(mat, fe.fe_newton_settings(defer_after=0, **kw), arr, U, xi0, outs)
    assert int(base["iters"].max()) >= 4
    for K in (1, None, 4):
        o = fe.fe_block_launch(mat, fe.fe_newton_settings(defer_after=K, **kw), arr, U, xi0, outs)
        torch.cuda.synchronize()
        for k in outs:      # same counts / flags; iterates, R_e, K_e of the two launches agree to rounding
            if k in ("iters", "flags"):     # (separately compiled instantiations: FMA contraction may differ)
                assert torch.equal(o[k], base[k]), (family, K, k)
            else:
                assert rel_err(o[k].cpu().numpy(), base[k].cpu().numpy()) < 1e-12, (family, K, k)


@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("kind", ["J2", "hosford"])
def test_mixed_block_jvp_and_vjp(cuda_device, family, kind):
    """K6 for the mixed u-p block: the JVP of both residual blocks (momentum rows through
    dev(d cauchy) - dp I, pressure rows through kappa, mu and the (u, p) direction) against
    the oracle (itself FD-checked in the CPU suite), K dU against the assembled mixed tangent,
    and the VJP through the adjoint identity over both blocks."""
    from cmad_b200 import Parameters, active_param_ids
    if kind == "J2":
        values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    else:
        values, act, tr = param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    pid = active_param_ids(P)
    nodes, conn = _mesh(family, (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn, mixed=True); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    rng = np.random.default_rng(11)
    n_e, n_ip, n_b = arr.n_elems, arr.n_ip, arr.n_basis
    nud = 3 * nodes.shape[0]
    stab = 0.7

    def Uvec(t, seed):
        U = np.zeros(arr.n_dofs)
        U[:nud] = fe_mesh.synthetic_displacement(nodes, t, seed=seed, ramp=0.004, noise=4e-4)
        U[nud:] = 30.0 * np.random.default_rng(seed).standard_normal(arr.n_dofs - nud)
        return torch.from_numpy(U).to(cuda_device)
    U1, U2 = Uvec(1.0, 1), Uvec(2.0, 2)
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    _, _, xi1 = fe.assemble_element_block_mixed(mat, nw, arr, U1, xi0, stab_mult=stab, want_K=False)
    _, vals, xi2 = fe.assemble_element_block_mixed(mat, nw, arr, U2, xi1, stab_mult=stab)
    assert float((xi2[..., 6] - xi1[..., 6]).max()) > 0.0                # plastic somewhere
    scale = np.array([3e3, 0.01, 1.5, 7.0, 4.0])
    dp = rng.standard_normal(len(pid)) * scale
    dxp = torch.from_numpy(1e-4 * rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    dUn = 1e-4 * rng.standard_normal(arr.n_dofs); dUn[nud:] *= 1e4
    dUv = torch.from_numpy(dUn).to(cuda_device)
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    eq, eqp = arr_h.elem_eq.numpy(), arr_h.elem_eq_p.numpy()
    geo = (arr_h.grad_N.numpy(), arr_h.N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy(), arr_h.h.numpy())
    args = (prob_eval, eq, eqp, U2.cpu().numpy(), xi1.cpu().numpy(), xi2.cpu().numpy()) + geo
    outs = {}
    for name, dxx, dUU in (("plain", dxp, None), ("nodxp", None, None), ("withU", dxp, dUv)):
        o = fe.fe_block_jvp(mat, arr, U2, xi1, xi2, pid, dp, dxx, outputs=("xi", "R_elem", "R_global"),
                            dU=dUU, stab_mult=stab)
        r = fe_oracle.block_jvp_mixed(*args, dp, None if dxx is None else dxx.cpu().numpy(), stab_mult=stab,
                                      dU=None if dUU is None else dUU.cpu().numpy())
        torch.cuda.synchronize()
        for k in ("xi", "R_elem", "R_p_elem"):
            assert rel_err(o[k].cpu().numpy(), r[k]) < 1e-9, (name, k, rel_err(o[k].cpu().numpy(), r[k]))
        Rg = np.zeros(arr.n_dofs)
        np.add.at(Rg, eq.reshape(-1), r["R_elem"].reshape(-1)); np.add.at(Rg, eqp.reshape(-1), r["R_p_elem"].reshape(-1))
        assert rel_err(o["R_global"].cpu().numpy(), Rg) < 1e-9
        outs[name] = o
    # the displacement/pressure direction adds K dU over all four blocks of the mixed tangent
    nu_, np_ = 3 * n_b, n_b
    sz = [n_e * nu_ * nu_, n_e * nu_ * np_, n_e * np_ * nu_, n_e * np_ * np_]
    of = np.concatenate([[0], np.cumsum(sz)])
    Kuu = vals[of[0]:of[1]].view(n_e, nu_, nu_); Kup = vals[of[1]:of[2]].view(n_e, nu_, np_)
    Kpu = vals[of[2]:of[3]].view(n_e, np_, nu_); Kpp = vals[of[3]:of[4]].view(n_e, np_, np_)
    du_e, dp_e = dUv[arr.elem_eq.long()], dUv[arr.elem_eq_p.long()]
    KdU_u = torch.einsum("eij,ej->ei", Kuu, du_e) + torch.einsum("eij,ej->ei", Kup, dp_e)
    KdU_p = torch.einsum("eij,ej->ei", Kpu, du_e) + torch.einsum("eij,ej->ei", Kpp, dp_e)
    assert rel_err((outs["withU"]["R_elem"] - outs["plain"]["R_elem"]).cpu().numpy(), KdU_u.cpu().numpy()) < 1e-8
    assert rel_err((outs["withU"]["R_p_elem"] - outs["plain"]["R_p_elem"]).cpu().numpy(), KdU_p.cpu().numpy()) < 1e-8
    # ---- VJP: adjoint identity against the oracle JVP, bit-reproducible
    Rbar = torch.from_numpy(rng.standard_normal(arr.n_dofs)).to(cuda_device)
    Rbar[nud:] *= 1e3                              # the pressure rows are ~1e-3 of the momentum rows
    xibar = torch.from_numpy(rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    pbar, xbp = fe.fe_block_vjp(mat, arr, U2, xi1, xi2, pid, Rbar, xibar, stab_mult=stab)
    pbar2, xbp2 = fe.fe_block_vjp(mat, arr, U2, xi1, xi2, pid, Rbar, xibar, stab_mult=stab)
    torch.cuda.synchronize()
    assert torch.equal(pbar, pbar2) and torch.equal(xbp, xbp2)
    Rb, xb = Rbar.cpu().numpy(), xibar.cpu().numpy()
    for trial in range(3):
        dpt = rng.standard_normal(len(pid)) * scale
        dxt = 1e-4 * rng.standard_normal((n_e, n_ip, 7))
        jv = fe_oracle.block_jvp_mixed(*args, dpt, dxt, stab_mult=stab)
        dR = np.zeros(arr.n_dofs)
        np.add.at(dR, eq.reshape(-1), jv["R_elem"].reshape(-1)); np.add.at(dR, eqp.reshape(-1), jv["R_p_elem"].reshape(-1))
        lhs = float(Rb @ dR + (xb * jv["xi"]).sum())
        rhs = float(pbar.cpu().numpy() @ dpt + (xbp.cpu().numpy() * dxt).sum())
        terms = abs(Rb[:nud] @ dR[:nud]) + abs(Rb[nud:] @ dR[nud:]) + abs((xb * jv["xi"]).sum())
        assert abs(lhs - rhs) < 1e-9 * terms, (trial, lhs, rhs)
    # the pressure rows really contribute: dropping them changes the elastic entries of pbar
    Rb0 = Rbar.clone(); Rb0[nud:] = 0.0
    pbar_u, _ = fe.fe_block_vjp(mat, arr, U2, xi1, xi2, pid, Rb0, xibar, stab_mult=stab)
    assert not torch.allclose(pbar_u[:2], pbar[:2], rtol=1e-6, atol=0.0)
    assert torch.allclose(pbar_u[2:], pbar[2:], rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("family", ["tet4", "hex8"])
@pytest.mark.parametrize("kind", ["J2", "hill-rotated", "hosford"])
def test_block_vjp_disp_is_the_transpose_of_the_displacement_direction(cuda_device, family, kind):
    """cmadx_fe_block_vjp_disp against the ORACLE's JVP with a displacement direction:
    <Ubar, dU> == <xibar, dxi(dU)> + <Rbar, dR(dU)> for random directions and cotangents
    (xibar only, Rbar only, both), deterministic through the segment-sum plan."""
    from cmad_b200 import Parameters, active_param_ids
    if kind == "J2":
        values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    elif kind == "hosford":
        values, act, tr = param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y"))
    else:
        values, act, tr = param_tree("hill", hill=(0.45, 0.55, 0.5, 1.4, 1.5, 1.6),
                                     active=("E", "nu", "D", "S", "Y"),
                                     rotation=rotation_matrix([1.0, 2.0, -0.5], 0.7))
    P = Parameters(values, act, tr)
    nodes, conn = _mesh(family, (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    rng = np.random.default_rng(13)
    n_e, n_ip = arr.n_elems, arr.n_ip
    U1 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 1.0, seed=1, ramp=0.004, noise=4e-4)).to(cuda_device)
    U2 = torch.from_numpy(fe_mesh.synthetic_displacement(nodes, 2.0, seed=2, ramp=0.004, noise=4e-4)).to(cuda_device)
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    xi1 = fe.fe_block_launch(mat, nw, arr, U1, xi0, ("xi",))["xi"]
    prim = fe.fe_block_launch(mat, nw, arr, U2, xi1, ("xi", "flags"))
    assert bool((prim["flags"] & 2).any())
    Rbar = torch.from_numpy(rng.standard_normal(arr.n_dofs)).to(cuda_device)
    xibar = torch.from_numpy(rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    plan = fe.disp_cotangent_plan(arr, device=cuda_device)
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    geo = (arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy())
    eq = arr_h.elem_eq.numpy()
    na = len(prob_eval.active_pid)
    for xb, rb in ((xibar, None), (None, Rbar), (xibar, Rbar)):
        ub_ip = fe.fe_block_vjp_disp(mat, arr, U2, xi1, prim["xi"], xb, rb)
        ub = plan.sum(ub_ip.reshape(-1))
        ub2 = plan.sum(fe.fe_block_vjp_disp(mat, arr, U2, xi1, prim["xi"], xb, rb).reshape(-1))
        torch.cuda.synchronize()
        assert torch.equal(ub, ub2)
        for trial in range(2):
            dU = 1e-4 * rng.standard_normal(arr.n_dofs)
            jv = fe_oracle.block_jvp(prob_eval, eq, U2.cpu().numpy(), xi1.cpu().numpy(), prim["xi"].cpu().numpy(),
                                     *geo, np.zeros(na), None, dU=dU)
            dR = np.zeros(arr.n_dofs); np.add.at(dR, eq.reshape(-1), jv["R_elem"].reshape(-1))
            t1 = float((xb.cpu().numpy() * jv["xi"]).sum()) if xb is not None else 0.0
            t2 = float(rb.cpu().numpy() @ dR) if rb is not None else 0.0
            lhs = float(ub.cpu().numpy() @ dU)
            assert abs(lhs - (t1 + t2)) < 1e-9 * (abs(t1) + abs(t2)), (trial, lhs, t1, t2)


@pytest.mark.parametrize("family,degree,n_ip", [("tet4", 2, 4), ("hex8", 4, 27), ("hex8", 1, 1)])
@pytest.mark.parametrize("kind", ["J2", "hosford"])
def test_block_parity_other_quadrature_rules(cuda_device, family, degree, n_ip, kind):
    """`discretization.quadrature.volume degree` overrides (cmad/cli/common.py:497-540): tet4 x 4
    (what the mixed formulation requires on tets), hex8 x 27, hex8 x 1 through the any-rule
    kernel (fe_generic.cu) against the block oracle: values 1e-10, counts and flags exact; then
    the K6 companions (JVP vs oracle, VJP adjoint identity) and evaluate_cauchy_at_ips."""
    from cmad_b200 import Parameters, active_param_ids
    values, act, tr = (param_tree("J2", active=("E", "nu", "D", "S", "Y")) if kind == "J2" else
                       param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y")))
    P = Parameters(values, act, tr)
    pid = active_param_ids(P)
    nodes, conn = _mesh(family, (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn, volume_degree=degree); arr = arr_h.to(cuda_device)
    assert arr.n_ip == n_ip
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    n_e = arr.n_elems
    geo = (arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy())
    eq = arr_h.elem_eq.numpy()
    xi_ref = np.zeros((n_e, n_ip, 7))
    xi = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    plastic = False
    for s in (1, 2):
        U = fe_mesh.synthetic_displacement(nodes, t=float(s), seed=40 + s, ramp=0.004, noise=4e-4)
        Ud = torch.from_numpy(U).to(cuda_device)
        xi_prev_d, xi_prev_ref = xi, xi_ref
        out = fe.fe_block_launch(mat, nw, arr, Ud, xi, outputs=("xi", "R_elem", "K_elem", "sigma", "iters", "flags"))
        ref = fe_oracle.assemble_block(prob, eq, U, xi_ref, *geo)
        torch.cuda.synchronize()
        assert np.array_equal(out["iters"].cpu().numpy(), ref["iters"])
        assert np.array_equal(out["flags"].cpu().numpy(), ref["flags"])
        for k in ("xi", "sigma", "R_elem", "K_elem"):
            assert rel_err(out[k].cpu().numpy(), ref[k]) < TOL, (k, s, rel_err(out[k].cpu().numpy(), ref[k]))
        # residual-only variant and the atomic global scatter
        o4 = fe.fe_block_launch(mat, nw, arr, Ud, xi, outputs=("xi", "R_elem", "R_global"))
        # K3 and K4 are separately compiled instantiations (FMA contraction may differ): rounding level
        assert rel_err(o4["R_elem"].cpu().numpy(), out["R_elem"].cpu().numpy()) < 1e-13
        Rg = np.zeros(arr.n_dofs); np.add.at(Rg, eq.reshape(-1), ref["R_elem"].reshape(-1))
        assert rel_err(o4["R_global"].cpu().numpy(), Rg) < 1e-9
        xi, xi_ref = out["xi"], ref["xi"]
        plastic |= bool((ref["flags"] & 2).any())
    assert plastic
    # evaluate_cauchy_at_ips at the stored state
    sig = fe.evaluate_cauchy_at_ips(mat, arr, Ud, xi)
    assert rel_err(sig.cpu().numpy(), ref["sigma"]) < TOL
    # K6: JVP against the oracle, VJP / displacement cotangent through the adjoint identity
    rng = np.random.default_rng(3)
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    dp = rng.standard_normal(len(pid)) * np.array([3e3, 0.01, 1.5, 7.0, 4.0])
    dxp = 1e-4 * rng.standard_normal((n_e, n_ip, 7))
    dU = 1e-4 * rng.standard_normal(arr.n_dofs)
    jv = fe.fe_block_jvp(mat, arr, Ud, xi_prev_d, xi, pid, dp, torch.from_numpy(dxp).to(cuda_device),
                         dU=torch.from_numpy(dU).to(cuda_device))
    jr = fe_oracle.block_jvp(prob_eval, eq, U, xi_prev_ref, xi_ref, *geo, dp, dxp, dU=dU)
    assert rel_err(jv["xi"].cpu().numpy(), jr["xi"]) < 1e-9 and rel_err(jv["R_elem"].cpu().numpy(), jr["R_elem"]) < 1e-9
    Rbar = torch.from_numpy(rng.standard_normal(arr.n_dofs)).to(cuda_device)
    xibar = torch.from_numpy(rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    pbar, xbp = fe.fe_block_vjp(mat, arr, Ud, xi_prev_d, xi, pid, Rbar, xibar)
    ub = fe.disp_cotangent_plan(arr, device=cuda_device).sum(
        fe.fe_block_vjp_disp(mat, arr, Ud, xi_prev_d, xi, xibar, Rbar).reshape(-1))
    dR = np.zeros(arr.n_dofs); np.add.at(dR, eq.reshape(-1), jr["R_elem"].reshape(-1))
    lhs = float(Rbar.cpu().numpy() @ dR + (xibar.cpu().numpy() * jr["xi"]).sum())
    rhs = float(pbar.cpu().numpy() @ dp + (xbp.cpu().numpy() * dxp).sum() + ub.cpu().numpy() @ dU)
    terms = abs(Rbar.cpu().numpy() @ dR) + abs((xibar.cpu().numpy() * jr["xi"]).sum())
    assert abs(lhs - rhs) < 1e-9 * terms, (lhs, rhs)


@pytest.mark.parametrize("kind", ["J2", "hosford"])
def test_mixed_tet4_with_the_degree_2_rule(cuda_device, kind):
    """The reference requires volume degree >= 2 for the mixed u-p formulation
    (cmad/cli/common.py:379-391): on tets that is the 4-point rule.  All six outputs of the mixed
    block, the K6-mixed JVP and the VJP adjoint identity against the oracle."""
    from cmad_b200 import Parameters, active_param_ids
    values, act, tr = (param_tree("J2", active=("E", "nu", "D", "S", "Y")) if kind == "J2" else
                       param_tree("hosford", a=6.0, active=("E", "nu", "D", "S", "Y")))
    P = Parameters(values, act, tr)
    pid = active_param_ids(P)
    nodes, conn = _mesh("tet4", (3, 2, 2))
    arr_h = fe_mesh.block_arrays(nodes, conn, mixed=True, volume_degree=2); arr = arr_h.to(cuda_device)
    n_e, n_ip, n_b = arr.n_elems, arr.n_ip, arr.n_basis
    assert (n_ip, n_b) == (4, 4)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    nud = 3 * nodes.shape[0]
    rng = np.random.default_rng(21)
    stab = 0.7

    def Uvec(t, seed):
        U = np.zeros(arr.n_dofs)
        U[:nud] = fe_mesh.synthetic_displacement(nodes, t, seed=seed, ramp=0.004, noise=4e-4)
        U[nud:] = 30.0 * np.random.default_rng(seed).standard_normal(arr.n_dofs - nud)
        return U
    eq, eqp = arr_h.elem_eq.numpy(), arr_h.elem_eq_p.numpy()
    geo = (arr_h.grad_N.numpy(), arr_h.N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy(), arr_h.h.numpy())
    U1, U2 = Uvec(1.0, 1), Uvec(2.0, 2)
    z = np.zeros((n_e, n_ip, 7))
    r1 = fe_oracle.assemble_block_mixed(prob, eq, eqp, U1, z, *geo, stab_mult=stab, want_K=False)
    r2 = fe_oracle.assemble_block_mixed(prob, eq, eqp, U2, r1["xi"], *geo, stab_mult=stab)
    U1d, U2d = (torch.from_numpy(u).to(cuda_device) for u in (U1, U2))
    xi0 = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    _, _, xi1 = fe.assemble_element_block_mixed(mat, nw, arr, U1d, xi0, stab_mult=stab, want_K=False)
    R, vals, xi2 = fe.assemble_element_block_mixed(mat, nw, arr, U2d, xi1, stab_mult=stab,
                                                   r_plan=fe.mixed_r_plan(arr, device=cuda_device))
    torch.cuda.synchronize()
    assert rel_err(xi2.cpu().numpy(), r2["xi"]) < TOL and rel_err(R.cpu().numpy(), r2["R"]) < TOL
    ref_vals = np.concatenate([r2[k].reshape(-1) for k in ("K_uu", "K_up", "K_pu", "K_pp")])
    off = 0
    for k in ("K_uu", "K_up", "K_pu", "K_pp"):
        m_ = r2[k].size
        assert rel_err(vals[off:off + m_].cpu().numpy(), r2[k].reshape(-1)) < TOL, k
        off += m_
    assert off == ref_vals.size and float((xi2[..., 6] - xi1[..., 6]).max()) > 0.0
    # K6-mixed
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    dp = rng.standard_normal(len(pid)) * np.array([3e3, 0.01, 1.5, 7.0, 4.0])
    dxp = 1e-4 * rng.standard_normal((n_e, n_ip, 7))
    dUn = 1e-4 * rng.standard_normal(arr.n_dofs); dUn[nud:] *= 1e4
    jv = fe.fe_block_jvp(mat, arr, U2d, xi1, xi2, pid, dp, torch.from_numpy(dxp).to(cuda_device),
                         dU=torch.from_numpy(dUn).to(cuda_device), stab_mult=stab)
    jr = fe_oracle.block_jvp_mixed(prob_eval, eq, eqp, U2, r1["xi"], r2["xi"], *geo, dp, dxp, stab_mult=stab, dU=dUn)
    for k in ("xi", "R_elem", "R_p_elem"):
        assert rel_err(jv[k].cpu().numpy(), jr[k]) < 1e-9, k
    Rbar = torch.from_numpy(rng.standard_normal(arr.n_dofs)).to(cuda_device)
    xibar = torch.from_numpy(rng.standard_normal((n_e, n_ip, 7))).to(cuda_device)
    pbar, xbp = fe.fe_block_vjp(mat, arr, U2d, xi1, xi2, pid, Rbar, xibar, stab_mult=stab)
    j0 = fe_oracle.block_jvp_mixed(prob_eval, eq, eqp, U2, r1["xi"], r2["xi"], *geo, dp, dxp, stab_mult=stab)
    dR = np.zeros(arr.n_dofs)
    np.add.at(dR, eq.reshape(-1), j0["R_elem"].reshape(-1)); np.add.at(dR, eqp.reshape(-1), j0["R_p_elem"].reshape(-1))
    lhs = float(Rbar.cpu().numpy() @ dR + (xibar.cpu().numpy() * j0["xi"]).sum())
    rhs = float(pbar.cpu().numpy() @ dp + (xbp.cpu().numpy() * dxp).sum())
    terms = abs(Rbar.cpu().numpy()[:nud] @ dR[:nud]) + abs(Rbar.cpu().numpy()[nud:] @ dR[nud:]) + \
        abs((xibar.cpu().numpy() * j0["xi"]).sum())
    assert abs(lhs - rhs) < 1e-9 * terms, (lhs, rhs)


@pytest.mark.parametrize("family,degree", [("tet4", 2), ("tet4", None), ("hex8", None)])
def test_softening_elements_take_the_hand_back_pass(cuda_device, family, degree):
    """Voce softening (S < 0): the J2 radial-return first pass hands elements back and the second
    pass re-solves them with the generic Newton - for tet4 x 4 that second pass is the any-rule
    kernel in list mode (fe_generic.cu).  States / residual / tangent still match the oracle
    (counts compared on the agreeing points: softening sits on the branch knife-edge)."""
    from cmad_b200 import mp
    values, _, _ = param_tree("J2")
    values["plastic"]["flow stress"]["hardening"]["voce"] = {"S": -80.0, "D": 40.0}
    nodes, conn = _mesh(family, (4, 3, 3))
    arr_h = fe_mesh.block_arrays(nodes, conn, volume_degree=degree); arr = arr_h.to(cuda_device)
    mat = material_from_values(values)
    nw = fe.fe_newton_settings(**NEWTON)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    n_e, n_ip = arr.n_elems, arr.n_ip
    geo = (arr_h.grad_N.numpy(), arr_h.det.numpy(), arr_h.quad_w.numpy())
    xi_ref = np.zeros((n_e, n_ip, 7))
    xi = torch.zeros((n_e, n_ip, 7), dtype=torch.float64, device=cuda_device)
    bailed = 0
    for s in (1, 2, 3):
        U = fe_mesh.synthetic_displacement(nodes, t=float(s), seed=60 + s, ramp=0.004, noise=6e-4)
        out = fe.fe_block_launch(mat, nw, arr, torch.from_numpy(U).to(cuda_device), xi,
                                 outputs=("xi", "R_elem", "K_elem", "iters", "flags"))
        bailed += mp.debug_bail_count()
        ref = fe_oracle.assemble_block(prob, arr_h.elem_eq.numpy(), U, xi_ref, *geo)
        same = (out["iters"].cpu().numpy() == ref["iters"]) & (out["flags"].cpu().numpy() == ref["flags"])
        assert same.mean() > 0.99
        ok = same.all(axis=1)                                  # elements whose points all agree
        for k in ("xi", "R_elem", "K_elem"):
            g, r = out[k].cpu().numpy()[ok], ref[k][ok]
            assert rel_err(g, r) < 1e-9, (k, s, rel_err(g, r))
        xi, xi_ref = out["xi"], ref["xi"]
    assert bailed > 0
