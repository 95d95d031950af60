"""CPU tests of the FE element-block oracle and the synthetic-mesh helpers: the
NumPy/C++ block oracle against the line-by-line torch-AD restatement, and the
reference's known answers KA3 (reference hex) for the per-element kernel."""
import numpy as np
import torch

from cmad_b200 import fe_mesh
from oracle import analytic, cmad_oracle as co, fe_oracle, oracle_c as oc

NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)


def _plastic_hex_U():
    """tests/fem/test_per_element_coupled.py:80-91 (_U_elem_plastic_loading)."""
    U = np.zeros((8, 3))
    U[[1, 2, 5, 6], 0] = 0.005
    U[[2, 3, 6, 7], 1] = 0.003
    U[[4, 5, 6, 7], 2] = 0.002
    return U


def _ref_hex_block():
    nodes, conn = fe_mesh.structured_hex_mesh((1, 1, 1), (2.0, 2.0, 2.0), origin=(-1.0, -1.0, -1.0))
    return nodes, conn, fe_mesh.block_arrays(nodes, conn)


def test_mesh_helpers_reference_hex_and_tets():
    nodes, conn, arr = _ref_hex_block()
    # reference hex: iso_jac = I, physical shapes = reference shapes, det = 1, weights sum to 8
    xi, w = fe_mesh.hex_quadrature_deg2()
    N, g = fe_mesh.hex_linear_shapes(xi)
    assert np.allclose(arr.grad_N[0].numpy(), g, atol=1e-15)
    assert np.allclose(arr.det.numpy(), 1.0) and np.isclose(w.sum(), 8.0)
    assert np.allclose(N.sum(axis=1), 1.0) and np.allclose(g.sum(axis=1), 0.0, atol=1e-16)
    assert arr.elem_eq.dtype == torch.int32 and tuple(arr.elem_eq.shape) == (1, 24)
    # 6-tet split: positive volumes adding up to the hex volume
    nodes, conn = fe_mesh.structured_hex_mesh((2, 3, 2), (1.0, 1.5, 0.7))
    tets = fe_mesh.split_hex_to_tets(conn)
    ta = fe_mesh.block_arrays(nodes, tets)
    vol = (ta.det * ta.quad_w[None, :]).sum(axis=1).numpy()
    assert (vol > 0).all() and np.isclose(vol.sum(), 1.0 * 1.5 * 0.7)
    assert np.allclose(vol, 1.0 * 1.5 * 0.7 / (12 * 6))
    # dof numbering eq = node*3 + comp, (basis, comp) order
    assert np.array_equal(ta.elem_eq[5].numpy().reshape(4, 3), tets[5][:, None] * 3 + np.arange(3))


def test_coo_dedup_matches_reference_recipe():
    """assembled_coo_dedup (cmad/fem/assembly.py:1026-1070) by its own recipe:
    lexsort + group boundaries."""
    nodes, conn = fe_mesh.structured_hex_mesh((2, 2, 1))
    eq = fe_mesh.block_arrays(nodes, fe_mesh.split_hex_to_tets(conn)).elem_eq.numpy()
    ur, uc, scatter = fe_mesh.coo_dedup(eq)
    rows, cols = fe_mesh.coo_pattern(eq)
    perm = np.lexsort((cols, rows))
    sr, sc = rows[perm], cols[perm]
    new = np.ones(len(rows), bool)
    new[1:] = (sr[1:] != sr[:-1]) | (sc[1:] != sc[:-1])
    seg = np.cumsum(new) - 1
    ref_scatter = np.empty(len(rows), np.int64)
    ref_scatter[perm] = seg
    assert np.array_equal(ur, sr[new]) and np.array_equal(uc, sc[new])
    assert np.array_equal(scatter, ref_scatter)
    assert np.array_equal(ur[scatter], rows) and np.array_equal(uc[scatter], cols)


def test_block_oracle_matches_torch_ad_element_ka3():
    """KA3 (tests/fem/test_per_element_coupled.py:174-386): reference hex, plastic
    loading: output shapes (8,3), (8,3,8,3), (8,7); per-IP ||C|| < 1e-10; at least
    one plastic IP; and the vectorised block oracle equals the torch-AD element."""
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, conn, arr = _ref_hex_block()
    U = _plastic_hex_U()
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    xi_prev = np.zeros((1, 8, 7))
    Ug = np.zeros((8, 3)); Ug[conn[0]] = U          # element-local -> global node numbering
    blk = fe_oracle.assemble_block(prob, arr.elem_eq.numpy(), Ug.reshape(-1), xi_prev,
                                   arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    params = co.to_torch_tree(values)
    spec = co.ModelSpec()
    R, K, xs, infos = co.coupled_element(params, U, U * 0, xi_prev[0], arr.grad_N[0].numpy(),
                                         arr.det[0].numpy(), arr.quad_w.numpy(), spec, NEWTON)
    assert R.shape == (8, 3) and K.shape == (8, 3, 8, 3) and xs.shape == (8, 7)
    assert np.abs(blk["R_elem"].reshape(8, 3) - R.numpy()).max() < 1e-10 * np.abs(R.numpy()).max()
    assert np.abs(blk["K_elem"].reshape(8, 3, 8, 3) - K.numpy()).max() < 1e-10 * np.abs(K.numpy()).max()
    assert np.abs(blk["xi"][0] - xs).max() < 1e-13
    assert [i.iters for i in infos] == list(blk["iters"][0])
    assert (blk["flags"][0] & 2).any() and (xs[:, 6] > 0).any()
    for ip in range(8):
        gu = co.interpolate_grad_u(torch.as_tensor(U), arr.grad_N[0, ip])
        C = co.sep_residual(torch.as_tensor(xs[ip]), torch.zeros(7, dtype=co.DT), params, gu, gu * 0, spec)
        assert float(torch.linalg.norm(C)) < 1e-10


def test_block_oracle_tangent_vs_central_fd_tets():
    """IFT-corrected K_e vs central differences of R_e (eps 1e-6, rtol 1e-5, atol
    1e-7 as tests/fem/test_per_element_coupled.py), on a small distorted tet block."""
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, conn = fe_mesh.structured_hex_mesh((1, 1, 1))
    rng = np.random.default_rng(3)
    nodes = nodes + 0.05 * rng.standard_normal(nodes.shape)
    arr = fe_mesh.block_arrays(nodes, fe_mesh.split_hex_to_tets(conn))
    U = fe_mesh.synthetic_displacement(nodes, t=2.0, seed=5, noise=5e-4)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    xi_prev = np.zeros((arr.n_elems, 1, 7))
    args = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    eq = arr.elem_eq.numpy()
    blk = fe_oracle.assemble_block(prob, eq, U, xi_prev, *args)
    assert (blk["flags"] & 2).any()
    eps = 1e-6
    for e in range(arr.n_elems):
        fd = np.zeros((12, 12))
        for c in range(12):
            Up, Um = U.copy(), U.copy()
            Up[eq[e, c]] += eps; Um[eq[e, c]] -= eps
            rp = fe_oracle.assemble_block(prob, eq[e:e + 1], Up, xi_prev[e:e + 1], args[0][e:e + 1], args[1][e:e + 1], args[2], want_K=False)
            rm = fe_oracle.assemble_block(prob, eq[e:e + 1], Um, xi_prev[e:e + 1], args[0][e:e + 1], args[1][e:e + 1], args[2], want_K=False)
            fd[:, c] = (rp["R_elem"][0] - rm["R_elem"][0]) / (2 * eps)
        assert np.allclose(blk["K_elem"][e], fd, rtol=1e-5, atol=1e-7 * np.abs(fd).max())
    # scatter-add of R and the COO dedup are plain segment sums
    ur, uc, scatter = fe_mesh.coo_dedup(eq)
    Kd = fe_oracle.coo_dedup_sum(blk["K_elem"].reshape(-1), scatter, len(ur))
    dense = np.zeros((len(U), len(U)))
    rows, cols = fe_mesh.coo_pattern(eq)
    np.add.at(dense, (rows, cols), blk["K_elem"].reshape(-1))
    assert np.allclose(dense[ur, uc], Kd, rtol=1e-13)
    assert np.isclose(blk["R"].sum(), blk["R_elem"].sum())


# ---- multi-rank host logic (gloo, world size 2): element partition + R all-reduce ----
def _fe_gloo_worker(rank, world, port, nodes, conn, U, ret):
    import os
    import torch.distributed as dist
    from cmad_b200 import fe
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    values, _, _ = analytic.j2_voce_param_tree("J2")
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    arr = fe_mesh.block_arrays(nodes, conn)
    local, (lo, hi) = fe.partition_block(arr, rank, world)          # product host logic
    o = fe_oracle.assemble_block(prob, local.elem_eq.numpy(), U, np.zeros((hi - lo, arr.n_ip, 7)),
                                 local.grad_N.numpy(), local.det.numpy(), local.quad_w.numpy())
    from cmad_b200.comm import WORLD
    R = fe.reduce_residual(torch.from_numpy(o["R"].copy()), WORLD)  # product exchange step (opt-in)
    R_none = fe.reduce_residual(torch.from_numpy(o["R"].copy()))    # group=None: no collective
    assert np.array_equal(R_none.numpy(), o["R"])
    # halo variant: only the dofs shared between ranks are exchanged
    halo = fe.InterfaceExchange(local.elem_eq, arr.n_dofs, group=WORLD)
    R_halo = halo.reduce(torch.from_numpy(o["R"].copy()))
    ret[rank] = (R.numpy(), lo, hi, o["K_elem"], R_halo.numpy(), halo.mine.numpy(), halo.n_interface)
    dist.destroy_process_group()


def test_two_rank_gloo_element_partition_equals_single_process():
    import os
    import torch.multiprocessing as tmp
    values, _, _ = analytic.j2_voce_param_tree("J2")
    nodes, conn = fe_mesh.structured_hex_mesh((3, 1, 1))            # 3 hexes: ragged 2 + 1 split
    U = fe_mesh.synthetic_displacement(nodes, t=2.0, seed=9, noise=3e-4)
    prob = oc.describe(values, None, newton_mode="traced", strain_comps=9, **NEWTON)
    arr = fe_mesh.block_arrays(nodes, conn)
    full = fe_oracle.assemble_block(prob, arr.elem_eq.numpy(), U, np.zeros((3, 8, 7)),
                                    arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    ret = tmp.Manager().dict()
    tmp.spawn(_fe_gloo_worker, args=(2, 29500 + os.getpid() % 2000, nodes, conn, U, ret), nprocs=2, join=True)
    for r in (0, 1):
        R, lo, hi, K, R_halo, mine, n_if = ret[r]
        assert np.allclose(R, full["R"], rtol=1e-13, atol=1e-13 * np.abs(full["R"]).max())
        assert np.array_equal(K, full["K_elem"][lo:hi])             # element-owned data stays local
        # interface exchange: complete on every dof this rank touches, 4 shared nodes x 3 dofs moved
        assert n_if == 12 and mine.sum() in (36, 24)
        assert np.allclose(R_halo[mine], full["R"][mine], rtol=1e-13, atol=1e-13 * np.abs(full["R"]).max())
    assert (ret[0][1], ret[0][2], ret[1][1], ret[1][2]) == (0, 2, 2, 3)


def test_block_jvp_oracle_vs_central_fd():
    """The K6 oracle (IFT sensitivities of xi and R_e w.r.t. parameters and xi_prev at
    fixed U) against central differences of the Newton-running primal block - the
    reference's FE FD-check pattern (tests/fem/test_fem_fd_checks.py:325-666)."""
    from cmad_b200 import Parameters
    from tests.helpers import param_tree
    values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    nodes, conn = fe_mesh.structured_hex_mesh((2, 1, 1))
    arr = fe_mesh.block_arrays(nodes, conn)
    U = fe_mesh.synthetic_displacement(nodes, t=2.0, seed=3, noise=4e-4)
    eq = arr.elem_eq.numpy(); geo = (arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
    tight = dict(max_iters=30, abs_tol=1e-14, rel_tol=1e-14)

    def primal(vals, xi_prev, Uv=None):
        prob = oc.describe(vals, P.active_idx, newton_mode="traced", strain_comps=9, **tight)
        return fe_oracle.assemble_block(prob, eq, U if Uv is None else Uv, xi_prev, *geo, want_K=False)

    rng = np.random.default_rng(0)
    xi_prev = primal(values, np.zeros((2, 8, 7)))["xi"] * 0.5          # a non-trivial previous state
    base = primal(values, xi_prev)
    assert (base["flags"] & 2).any()
    dp = np.array([3.0e3, 0.01, 1.5, -7.0, 4.0])                       # E, nu, D, S, Y (native)
    dxp = 1e-4 * rng.standard_normal(xi_prev.shape)
    dxp[:, :, 6] = np.abs(dxp[:, :, 6])
    prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)
    dUv = 1e-4 * rng.standard_normal(U.shape)

    import copy
    def shifted(h, with_U):
        v = copy.deepcopy(values)
        v["elastic"]["E"] += h * dp[0]; v["elastic"]["nu"] += h * dp[1]
        vo = v["plastic"]["flow stress"]["hardening"]["voce"]
        vo["D"] += h * dp[2]; vo["S"] += h * dp[3]
        v["plastic"]["flow stress"]["initial yield"]["Y"] += h * dp[4]
        return primal(v, xi_prev + h * dxp, U + h * dUv if with_U else None)
    h = 1e-5
    for with_U in (False, True):
        jv = fe_oracle.block_jvp(prob_eval, eq, U, xi_prev, base["xi"], *geo, dp, dxp,
                                 dU=dUv if with_U else None)
        up, dn = shifted(h, with_U), shifted(-h, with_U)
        fd_R = (up["R_elem"] - dn["R_elem"]) / (2 * h)
        fd_x = (up["xi"] - dn["xi"]) / (2 * h)
        assert np.abs(jv["R_elem"] - fd_R).max() < 1e-6 * np.abs(fd_R).max()
        assert np.abs(jv["xi"] - fd_x).max() < 1e-6 * np.abs(fd_x).max()


def test_block_jvp_mixed_oracle_vs_central_fd():
    """The mixed u-p K6 oracle (both residual blocks, parameters entering R_p through kappa
    and mu) against central differences of the Newton-running mixed primal block."""
    from cmad_b200 import Parameters
    from tests.helpers import param_tree
    import copy
    values, act, tr = param_tree("J2", active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    nodes, conn = fe_mesh.structured_hex_mesh((2, 1, 1))
    rng = np.random.default_rng(5)
    for family in ("hex8", "tet4"):
        cn = conn if family == "hex8" else fe_mesh.split_hex_to_tets(conn)
        arr = fe_mesh.block_arrays(nodes, cn, mixed=True)
        n_e, n_ip = arr.n_elems, arr.n_ip
        nu_dofs = 3 * nodes.shape[0]
        U = np.zeros(arr.n_dofs)
        U[:nu_dofs] = fe_mesh.synthetic_displacement(nodes, t=2.0, seed=3, noise=4e-4)
        U[nu_dofs:] = 30.0 * rng.standard_normal(arr.n_dofs - nu_dofs)
        eq, eqp = arr.elem_eq.numpy(), arr.elem_eq_p.numpy()
        geo = (arr.grad_N.numpy(), arr.N.numpy(), arr.det.numpy(), arr.quad_w.numpy(), arr.h.numpy())
        tight = dict(max_iters=30, abs_tol=1e-14, rel_tol=1e-14)

        def primal(vals, xi_prev, Uv):
            prob = oc.describe(vals, P.active_idx, newton_mode="traced", strain_comps=9, **tight)
            return fe_oracle.assemble_block_mixed(prob, eq, eqp, Uv, xi_prev, *geo, stab_mult=0.7, want_K=False)

        xi_prev = primal(values, np.zeros((n_e, n_ip, 7)), U)["xi"] * 0.5
        base = primal(values, xi_prev, U)
        dp = np.array([3.0e3, 0.01, 1.5, -7.0, 4.0])
        dxp = 1e-4 * rng.standard_normal(xi_prev.shape)
        dxp[:, :, 6] = np.abs(dxp[:, :, 6])
        dUv = 1e-4 * rng.standard_normal(U.shape)
        dUv[nu_dofs:] *= 1e4
        prob_eval = oc.describe(values, P.active_idx, newton_mode="imperative", strain_comps=9, max_iters=0)

        def shifted(hh, with_U):
            v = copy.deepcopy(values)
            v["elastic"]["E"] += hh * dp[0]; v["elastic"]["nu"] += hh * dp[1]
            vo = v["plastic"]["flow stress"]["hardening"]["voce"]
            vo["D"] += hh * dp[2]; vo["S"] += hh * dp[3]
            v["plastic"]["flow stress"]["initial yield"]["Y"] += hh * dp[4]
            return primal(v, xi_prev + hh * dxp, U + hh * dUv if with_U else U)
        hh = 1e-5
        for with_U in (False, True):
            jv = fe_oracle.block_jvp_mixed(prob_eval, eq, eqp, U, xi_prev, base["xi"], *geo, dp, dxp,
                                           stab_mult=0.7, dU=dUv if with_U else None)
            up, dn = shifted(hh, with_U), shifted(-hh, with_U)
            for k, ku in (("R_elem", "R_u"), ("R_p_elem", "R_p"), ("xi", "xi")):
                fd = (up[ku] - dn[ku]) / (2 * hh)
                assert np.abs(jv[k] - fd).max() < 2e-6 * np.abs(fd).max(), (family, with_U, k)
