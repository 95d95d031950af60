"""Block / global FE layer against golden vectors produced by EXECUTING THE REFERENCE'S OWN
SOURCE (tests/golden/make_reference_fe_block_golden.py): `build_fe_kernel_arrays`,
`assemble_element_block`, `assemble_global` + `assembled_coo_dedup`, `_embedded_bc_enforce` /
`_embedded_residual` and `fe_quasistatic_drive` on a `StructuredHexMesh` and its
`hex_to_tet_split`, displacement and mixed u-p.  Fixture: tests/golden/ref_fe_block.npz.

Integer layouts (equation gathers, COO pattern, dedup scatter, prescribed dofs, Newton counts)
are compared bit-exactly; floating-point results to 1e-10 relative (the arithmetic library of
the fixture run is NumPy/LAPACK instead of XLA, the algorithm is the reference's).

CPU: the host-side mesh / pattern builders and the oracle.  GPU: K3 / K3-mixed + K5 + the
embedded-BC kernels + the quasi-static driver over them."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cmad_b200 import fe_driver as drv, fe_mesh
from oracle import fe_oracle, oracle_c as oc
from tests.golden.materials import material

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fe_block.npz"))
CASES = sorted({".".join(k.split(".")[:2]) for k in G.files})
LOCAL_NEWTON = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
RAMP = 0.003


def g(case, key):
    return G[f"{case}.{key}"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(float(np.abs(b).max()), 1e-300))


def build(case):
    """This repo's host-side arrays for the fixture's mesh."""
    family, form = case.split(".")
    mixed = form == "mixed"
    nodes, conn = fe_mesh.structured_hex_mesh(tuple(int(d) for d in g(case, "divisions")))
    if family == "tet4":
        conn = fe_mesh.split_hex_to_tets(conn)
    n_ip = g(case, "quad_w").shape[0]
    degree = None
    if family == "tet4" and n_ip != 1:
        degree = 2
    arr = fe_mesh.block_arrays(nodes, conn, mixed=mixed, volume_degree=degree)
    return nodes, conn, arr, mixed


def bcs_of(case, nodes, n_u):
    """The uniaxial BCs of examples/elastic_plastic_uniaxial.yaml:58-66 in this repo's form,
    ordered as the fixture's `prescribed_indices`."""
    idx = g(case, "prescribed_indices").astype(np.int64)
    nid, comp = idx // 3, idx % 3
    ramped = (comp == 0) & np.isclose(nodes[nid, 0], 1.0)
    return drv.DirichletBCs(idx, lambda t: np.where(ramped, RAMP * t, 0.0))


def test_fixture_has_all_cases():
    assert CASES == ["hex8.disp", "hex8.mixed", "tet4.disp", "tet4.mixed"]


@pytest.mark.parametrize("case", CASES)
def test_mesh_and_kernel_arrays_equal_the_reference(case):
    nodes, conn, arr, mixed = build(case)
    assert np.array_equal(conn, g(case, "connectivity"))
    assert np.abs(nodes - g(case, "nodes")).max() < 1e-15
    n_e, n_b = conn.shape
    # u_gather_eq_by_block[block][field] (cmad/fem/kernel_arrays.py:68-72) / r_scatter_eq
    assert np.array_equal(arr.elem_eq.numpy().reshape(n_e, n_b, 3), g(case, "u_gather_eq.0"))
    assert np.array_equal(arr.elem_eq.numpy(), g(case, "r_scatter_eq.0"))
    assert arr.n_dofs == int(g(case, "n_dofs"))
    if mixed:
        assert np.array_equal(arr.elem_eq_p.numpy().reshape(n_e, n_b, 1), g(case, "u_gather_eq.1"))
        assert np.array_equal(arr.elem_eq_p.numpy(), g(case, "r_scatter_eq.1"))
        assert rel(arr.h.numpy(), g(case, "element_size")) < 1e-14
        assert rel(arr.grad_N.numpy(), g(case, "grad_N_phys.1")) < 1e-13
    assert rel(arr.quad_w.numpy(), g(case, "quad_w")) < 1e-15
    assert rel(arr.N.numpy(), g(case, "N.0")) < 1e-15
    assert rel(arr.det.numpy(), g(case, "iso_jac_det")) < 1e-13
    assert rel(arr.grad_N.numpy(), g(case, "grad_N_phys.0")) < 1e-13
    # assembled_coo_dedup (cmad/fem/assembly.py:1026-1070): integers, bit-exact
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    assert np.array_equal(ur, g(case, "coo_rows"))
    assert np.array_equal(uc, g(case, "coo_cols"))
    assert np.array_equal(scatter, g(case, "coo_dedup_scatter"))
    assert np.array_equal(np.stack([ur, uc], axis=-1), g(case, "asm0.K_indices"))


def oracle_assemble(case, arr, mixed, U, xi_prev):
    prob = oc.describe(material("J2"), None, newton_mode="traced", strain_comps=9, **LOCAL_NEWTON)
    eq = arr.elem_eq.numpy()
    if not mixed:
        o = fe_oracle.assemble_block(prob, eq, U, xi_prev, arr.grad_N.numpy(), arr.det.numpy(), arr.quad_w.numpy())
        return o["R"], o["K_elem"].reshape(-1), o["xi"]
    o = fe_oracle.assemble_block_mixed(prob, eq, arr.elem_eq_p.numpy(), U, xi_prev, arr.grad_N.numpy(),
                                       arr.N.numpy(), arr.det.numpy(), arr.quad_w.numpy(), arr.h.numpy(), 1.0)
    vals = np.concatenate([o[k].reshape(-1) for k in ("K_uu", "K_up", "K_pu", "K_pp")])
    return o["R"], vals, o["xi"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_block_and_global_assembly_vs_reference(case):
    """assemble_element_block / assemble_global / embedded BCs of the CPU oracle equal the
    reference's own run (this pins fe_oracle.embedded_system and the COO stream order)."""
    nodes, conn, arr, mixed = build(case)
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    for s in range(2):
        U, xi_prev = g(case, f"asm{s}.U"), g(case, f"asm{s}.xi_prev")
        R, vals, xi = oracle_assemble(case, arr, mixed, U, xi_prev)
        assert rel(xi, g(case, f"asm{s}.xi")) < 1e-10
        assert rel(R, g(case, f"asm{s}.R_block")) < 1e-10
        assert rel(R, g(case, f"asm{s}.R")) < 1e-10               # no Neumann terms: R_global == R_block
        assert rel(vals, g(case, f"asm{s}.vals")) < 1e-10         # with-duplicates stream, emit order
        K_data = fe_oracle.coo_dedup_sum(vals, scatter, len(ur))
        assert rel(K_data, g(case, f"asm{s}.K_data")) < 1e-10
        idx, pv = g(case, "prescribed_indices"), g(case, f"asm{s}.presc_vals")
        r, K_emb = fe_oracle.embedded_system(ur, uc, arr.n_dofs, K_data, R, U, idx, pv)
        assert rel(r, g(case, f"asm{s}.r_emb")) < 1e-10
        # the reference's K_emb = zeroed data on the pattern + appended (i, i, K_ii) entries
        ref = sp.coo_matrix((g(case, f"asm{s}.K_emb_data"), (np.concatenate([ur, idx]), np.concatenate([uc, idx]))),
                            shape=(arr.n_dofs,) * 2).toarray()
        mine = sp.coo_matrix((K_emb, (ur, uc)), shape=(arr.n_dofs,) * 2).toarray()
        assert rel(mine, ref) < 1e-10
        # ... and the host driver's own embedded_system (the one fe_newton_solve uses)
        bcs = drv.DirichletBCs(idx.astype(np.int64), lambda t, pv=pv: pv)
        r2, K2 = drv.embedded_system(drv.SparsePattern(ur, uc, arr.n_dofs), K_data, R, U, bcs, float(g(case, f"asm{s}.t")))
        assert rel(r2, g(case, f"asm{s}.r_emb")) < 1e-10
        assert rel(K2.toarray(), ref) < 1e-10


def drive(case, arr, mixed, nodes, assemble, xi0):
    ur, uc, _ = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    pattern = drv.SparsePattern(ur, uc, arr.n_dofs)
    bcs = bcs_of(case, nodes, nodes.shape[0] * 3)
    return drv.fe_quasistatic_drive(assemble, pattern, bcs, np.zeros(arr.n_dofs), xi0, g(case, "drive.t"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_quasistatic_drive_vs_reference(case):
    """fe_quasistatic_drive (cmad/fem/driver.py:149-253) with the deck defaults: per-step U,
    final xi, and the global Newton iteration count of every load step."""
    nodes, conn, arr, mixed = build(case)
    _, _, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    n_unique = len(g(case, "coo_rows"))

    def assemble(U, xi_prev):
        R, vals, xi = oracle_assemble(case, arr, mixed, U, xi_prev)
        return R, fe_oracle.coo_dedup_sum(vals, scatter, n_unique), xi
    U_steps, xi, _, logs = drive(case, arr, mixed, nodes, assemble, np.zeros((arr.n_elems, arr.n_ip, 7)))
    assert [l.iters for l in logs] == list(g(case, "drive.newton_iters"))
    Uref = g(case, "drive.U")[1:]
    assert rel(U_steps, Uref) < 1e-9
    assert rel(xi, g(case, "drive.xi")[-1]) < 1e-9


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("deterministic", [True, False])
def test_cuda_block_assembly_vs_reference(cuda_device, case, deterministic):
    """K3 / K3-mixed (+ K5) through the C-ABI vs the reference's assemble_element_block /
    assemble_global: R_block, the COO value stream in emit order, xi_solved, deduplicated K."""
    import torch
    from cmad_b200 import fe, material_from_values
    nodes, conn, arr, mixed = build(case)
    mat = material_from_values(material("J2"))
    nw = fe.fe_newton_settings(**LOCAL_NEWTON)
    arr_d = arr.to(cuda_device)
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    k_plan = fe.SegmentPlan(scatter, len(ur), device=cuda_device)
    if mixed:
        r_plan = fe.mixed_r_plan(arr_d, device=cuda_device) if deterministic else None
    else:
        r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device) if deterministic else None
    for s in range(2):
        U = torch.from_numpy(g(case, f"asm{s}.U")).to(cuda_device)
        xi_prev = torch.from_numpy(g(case, f"asm{s}.xi_prev")).to(cuda_device).contiguous()
        if mixed:
            R, vals, xi = fe.assemble_element_block_mixed(mat, nw, arr_d, U, xi_prev, stab_mult=1.0, r_plan=r_plan)
        else:
            R, vals, xi = fe.assemble_element_block(mat, nw, arr_d, U, xi_prev, r_plan=r_plan)
        K_data = k_plan.sum(vals)
        torch.cuda.synchronize()
        assert rel(xi.cpu().numpy(), g(case, f"asm{s}.xi")) < 1e-10
        assert rel(R.cpu().numpy(), g(case, f"asm{s}.R_block")) < 1e-10
        assert rel(vals.cpu().numpy(), g(case, f"asm{s}.vals")) < 1e-10
        assert rel(K_data.cpu().numpy(), g(case, f"asm{s}.K_data")) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_embedded_bcs_and_drive_vs_reference(cuda_device, case):
    """The device embedded-BC kernels (fe_post.cu) on the reference's assembled (K, R), and the
    quasi-static driver over the CUDA assembler: per-step U, xi and global Newton counts of the
    reference's own fe_quasistatic_drive run."""
    import torch
    from cmad_b200 import fe, material_from_values
    nodes, conn, arr, mixed = build(case)
    mat = material_from_values(material("J2"))
    nw = fe.fe_newton_settings(**LOCAL_NEWTON)
    arr_d = arr.to(cuda_device)
    ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.numpy(), arr.elem_eq_p.numpy() if mixed else None)
    pattern = drv.SparsePattern(ur, uc, arr.n_dofs)
    k_plan = fe.SegmentPlan(scatter, len(ur), device=cuda_device)
    if mixed:
        asm = drv.cuda_assembler_mixed(mat, nw, arr_d, fe.mixed_r_plan(arr_d, device=cuda_device), k_plan, 1.0)
    else:
        r_plan = fe.SegmentPlan(arr.elem_eq.numpy().reshape(-1), arr.n_dofs, device=cuda_device)
        asm = drv.cuda_assembler(mat, nw, arr_d, r_plan, k_plan)
    bcs = bcs_of(case, nodes, nodes.shape[0] * 3)
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=cuda_device)
    for embedded_on_device in (False, True):
        a = drv.DeviceEmbeddedBCs(asm, pattern, bcs, cuda_device) if embedded_on_device else asm
        if embedded_on_device:
            for s in range(2):      # the device embedded system at the fixture's assembly states
                U = g(case, f"asm{s}.U")
                xi_prev = torch.from_numpy(g(case, f"asm{s}.xi_prev")).to(cuda_device).contiguous()
                r, K_emb, _ = a.enforced(U, xi_prev, float(g(case, f"asm{s}.t")))
                idx = g(case, "prescribed_indices")
                ref = sp.coo_matrix((g(case, f"asm{s}.K_emb_data"),
                                     (np.concatenate([ur, idx]), np.concatenate([uc, idx]))),
                                    shape=(arr.n_dofs,) * 2).toarray()
                assert rel(np.asarray(r), g(case, f"asm{s}.r_emb")) < 1e-10
                assert rel(K_emb.toarray(), ref) < 1e-10
        U_steps, xi, _, logs = drv.fe_quasistatic_drive(a, pattern, bcs, np.zeros(arr.n_dofs), xi0, g(case, "drive.t"))
        assert [l.iters for l in logs] == list(g(case, "drive.newton_iters"))
        assert rel(U_steps, g(case, "drive.U")[1:]) < 1e-9
        assert rel(xi.cpu().numpy(), g(case, "drive.xi")[-1]) < 1e-9
