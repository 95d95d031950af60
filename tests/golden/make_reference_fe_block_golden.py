"""Golden vectors of the BLOCK / GLOBAL FE layer, produced by EXECUTING THE REFERENCE'S
OWN, UNMODIFIED SOURCE (`/root/reference/cmad/fem/...`) on the NumPy `jax` stand-in of
`tests/golden/jaxshim/` (see `make_reference_golden.py` for what the stand-in is).

What runs, as written in the reference, on a `StructuredHexMesh` (and its
`hex_to_tet_split`), displacement and mixed u-p formulations:

  * `build_dof_map`, `build_fe_problem`, `build_fe_kernel_arrays`
    (cmad/fem/kernel_arrays.py:58-228): `u_gather_eq_by_block`, `r_scatter_eq_by_block`,
    the deduplicated COO pattern and `coo_dedup_scatter`, the geometry cache, the
    prescribed dofs;
  * `assemble_element_block` (cmad/fem/assembly.py:616-732): `R_block`, the
    with-duplicates COO `vals` stream, `xi_solved`;
  * `assemble_global` + `assembled_coo_dedup` (assembly.py:816-917, 1026-1070);
  * `_embedded_bc_enforce` / `_embedded_residual` (cmad/fem/sparse_solve.py:1058-1176);
  * `fe_quasistatic_drive` (cmad/fem/driver.py:149-253) with the parameter sets and BCs of
    examples/elastic_plastic_uniaxial.yaml and examples/mixed_plastic.yaml: per-step U,
    xi, and the global Newton iteration counts (read off the stand-in's while-loop log).

Run in the build container only (needs /root/reference):

    python tests/golden/make_reference_fe_block_golden.py [--jobs 8]

Writes tests/golden/ref_fe_block.npz; committed, `/root/reference` is never read at test time.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def _enter_reference():
    import numpy as np
    sys.path[:0] = [os.path.join(HERE, "jaxshim"), REFERENCE, ROOT, HERE]
    for m in ("netCDF4", "gmsh", "matplotlib", "matplotlib.pyplot", "sympy", "jsonschema"):
        try:
            __import__(m)
        except ImportError:
            sys.modules[m] = types.ModuleType(m)
    try:
        import pyamg  # noqa: F401
    except ImportError:
        # only `near_null_space` (AMG preconditioner set-up, not on the path) needs pyamg:
        # the six rigid-body modes per node, interleaved dof order
        def coord_to_rbm(n, ndim, x, y, z):
            B = np.zeros((3 * n, 6))
            for k in range(3):
                B[k::3, k] = 1.0
            B[0::3, 3] = -y; B[1::3, 3] = x
            B[1::3, 4] = -z; B[2::3, 4] = y
            B[0::3, 5] = z; B[2::3, 5] = -x
            return B
        pa, pu, puu = (types.ModuleType(n) for n in ("pyamg", "pyamg.util", "pyamg.util.utils"))
        puu.coord_to_rbm = coord_to_rbm
        pa.util, pu.utils = pu, puu
        sys.modules.update({"pyamg": pa, "pyamg.util": pu, "pyamg.util.utils": puu})


_enter_reference()

import numpy as np  # noqa: E402

import jax.numpy as jnp  # noqa: E402  (the shim)
from jax import _core  # noqa: E402
from cmad.fem.assembly import assemble_element_block, assemble_global, params_by_block_from_models  # noqa: E402
from cmad.fem.bcs import DirichletBC  # noqa: E402
from cmad.fem.dof import GlobalFieldLayout, build_dof_map  # noqa: E402
from cmad.fem.driver import fe_quasistatic_drive  # noqa: E402
from cmad.fem.fe_problem import build_fe_problem  # noqa: E402
from cmad.fem.finite_element import P1_TET, Q1_HEX  # noqa: E402
from cmad.fem.mesh import StructuredHexMesh, hex_to_tet_split  # noqa: E402
from cmad.fem.sparse_solve import _embedded_bc_enforce, _embedded_residual  # noqa: E402
from cmad.global_residuals.modes import GlobalResidualMode  # noqa: E402
from cmad.global_residuals.small_disp_equilibrium import SmallDispEquilibrium  # noqa: E402
from cmad.models.small_elastic_plastic import SmallElasticPlastic  # noqa: E402
from cmad.parameters.parameters import Parameters  # noqa: E402

from materials import active_all_scalars, const_like, material  # noqa: E402

RAMP = 0.003           # examples/elastic_plastic_uniaxial.yaml:66 (3 x yield strain at t = 1)


def _dbcs():
    """examples/elastic_plastic_uniaxial.yaml:58-66 / mixed_plastic.yaml: symmetry pins on the
    three min faces, u_x = 0.003 t on the +x face."""
    def ux(coords, t):
        return jnp.full((np.asarray(coords).shape[0], 1), RAMP * t)
    return [DirichletBC(sideset_names=["xmin_sides"], field_name="u", dofs=(0,), values=None),
            DirichletBC(sideset_names=["ymin_sides"], field_name="u", dofs=(1,), values=None),
            DirichletBC(sideset_names=["zmin_sides"], field_name="u", dofs=(2,), values=None),
            DirichletBC(sideset_names=["xmax_sides"], field_name="u", dofs=(0,), values=ux)]


def _problem(family, mixed, divisions):
    mesh = StructuredHexMesh(lengths=(1.0, 1.0, 1.0), divisions=divisions)
    fe = Q1_HEX
    if family == "tet4":
        mesh = hex_to_tet_split(mesh)
        fe = P1_TET
    layouts = [GlobalFieldLayout(name="u", finite_element=fe)]
    comps = {"u": 3}
    if mixed:
        layouts.append(GlobalFieldLayout(name="p", finite_element=fe))
        comps["p"] = 1
    dof_map = build_dof_map(mesh, layouts, _dbcs(), components_by_field=comps)
    values = material("J2")                  # E 200e3, nu 0.3, Y 200, Voce S 200 / D 20 (both decks)
    P = Parameters(values, active_all_scalars(values), const_like(values, None))
    model = SmallElasticPlastic(P)
    gr = SmallDispEquilibrium(ndims=3, mixed=mixed) if mixed else SmallDispEquilibrium(ndims=3)
    fp = build_fe_problem(mesh=mesh, dof_map=dof_map, gr=gr, models_by_block={"all": model},
                          modes_by_block={"all": GlobalResidualMode.COUPLED})
    return mesh, fp


def _case(job):
    family, mixed, divisions, n_steps = job
    mesh, fp = _problem(family, mixed, divisions)
    ka = fp.kernel_arrays
    geom = ka.geometry_cache["all"]
    out = {"nodes": np.asarray(mesh.nodes), "connectivity": np.asarray(mesh.connectivity),
           "divisions": np.asarray(divisions), "n_dofs": np.asarray(fp.dof_map.num_total_dofs),
           "block_offsets": np.asarray(fp.dof_map.block_offsets),
           "coo_rows": np.asarray(ka.coo_rows), "coo_cols": np.asarray(ka.coo_cols),
           "coo_dedup_scatter": np.asarray(ka.coo_dedup_scatter),
           "prescribed_indices": np.asarray(ka.prescribed_indices),
           "quad_w": np.asarray(geom.shared.quad_w),
           "iso_jac_det": np.asarray(geom.per_elem.iso_jac_det),
           "element_size": np.asarray(geom.per_elem.element_size)}
    for f, eq in enumerate(ka.u_gather_eq_by_block["all"]):
        out[f"u_gather_eq.{f}"] = np.asarray(eq)
    for r, eq in enumerate(ka.r_scatter_eq_by_block["all"]):
        out[f"r_scatter_eq.{r}"] = np.asarray(eq)
    for r, g in enumerate(geom.per_elem.field_grad_N_phys_per_block):
        out[f"grad_N_phys.{r}"] = np.asarray(g)
    for r, N in enumerate(geom.shared.field_N_per_block):
        out[f"N.{r}"] = np.asarray(N)

    # ---- one block / global assembly at a random state (two states: virgin and hardened xi_prev)
    pb = params_by_block_from_models(fp)
    n = int(fp.dof_map.num_total_dofs)
    n_u = int(fp.dof_map.block_offsets[1]) if mixed else n
    n_e = mesh.connectivity.shape[0]
    n_ip = int(np.asarray(geom.shared.quad_w).shape[0])
    rng = np.random.default_rng(11 + 7 * len(family) + (3 if mixed else 0))
    xi_prev = np.zeros((n_e, n_ip, 7))
    presc_idx = ka.prescribed_indices
    for s in range(2):
        U = np.zeros(n)
        x = np.asarray(mesh.nodes)
        U[0:n_u:3] = (0.002 + 0.0015 * s) * x[:, 0]
        U[:n_u] += 3e-4 * rng.standard_normal(n_u)
        if mixed:
            U[n_u:] = -60.0 + 25.0 * rng.standard_normal(n - n_u)
        t = 0.5 * (s + 1)
        R_block, vals, xi = assemble_element_block(fp, ka, pb, "all", U, U, t=t, xi_prev_per_block=xi_prev)
        K, R, xis = assemble_global(fp, ka, pb, U, U, t=t, xi_prev_by_block={"all": xi_prev})
        presc_vals = jnp.asarray(fp.dof_map.evaluate_prescribed_values(ka.dbc_arrays, t))
        K_emb, K_ii = _embedded_bc_enforce(K, presc_idx)
        r_emb = _embedded_residual(R, K, jnp.asarray(U), presc_idx, presc_vals, K_ii)
        out.update({f"asm{s}.U": U, f"asm{s}.t": np.asarray(t), f"asm{s}.xi_prev": xi_prev.copy(),
                    f"asm{s}.R_block": np.asarray(R_block), f"asm{s}.vals": np.asarray(vals),
                    f"asm{s}.xi": np.asarray(xi), f"asm{s}.K_data": np.asarray(K.data),
                    f"asm{s}.K_indices": np.asarray(K.indices), f"asm{s}.R": np.asarray(R),
                    f"asm{s}.presc_vals": np.asarray(presc_vals),
                    f"asm{s}.K_emb_data": np.asarray(K_emb), f"asm{s}.K_ii_presc": np.asarray(K_ii),
                    f"asm{s}.r_emb": np.asarray(r_emb)})
        xi_prev = np.asarray(xi)

    # ---- the quasi-static drive (deck defaults: Newton 10 / 1e-10 / 1e-10, cubic line search)
    _core.WHILE_LOG.clear()
    ts = list(np.linspace(0.0, 1.0, n_steps + 1))
    state, J = fe_quasistatic_drive(fp, ts)
    iters = [int(cnt) for cnt, carry in _core.WHILE_LOG
             if isinstance(carry, tuple) and len(carry) == 5 and isinstance(carry[4], dict)]
    out["drive.t"] = np.asarray(ts)
    out["drive.U"] = np.array([np.asarray(state.U_at(k)) for k in range(len(ts))])
    out["drive.xi"] = np.array([np.asarray(state.xi_at(k, "all")) for k in range(len(ts))])
    out["drive.newton_iters"] = np.asarray(iters)
    return out


CASES = [("hex8", False, (2, 2, 1), 3), ("hex8", True, (2, 1, 1), 2),
         ("tet4", False, (1, 1, 1), 3), ("tet4", True, (1, 1, 1), 2)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    cases = [c for c in CASES if not a.only or f"{c[0]}.{'mixed' if c[1] else 'disp'}" in a.only.split(",")]
    with mp.Pool(min(a.jobs, len(cases))) as pool:
        res = pool.map(_case, cases, chunksize=1)
    path = os.path.join(HERE, "ref_fe_block.npz")
    out = dict(np.load(path)) if (a.only and os.path.exists(path)) else {}
    for c, r in zip(cases, res):
        name = f"{c[0]}.{'mixed' if c[1] else 'disp'}"
        for k, v in r.items():
            out[f"{name}.{k}"] = v
        print(name, "dofs", int(r["n_dofs"]), "nnz", r["coo_rows"].shape[0], "newton iters", r["drive.newton_iters"],
              "alpha max", r["drive.xi"][-1][..., 6].max(), flush=True)
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
