"""YAML deck front end for the FE hot path: the reference's own decks (examples/*.yaml) are
read as they are and turned into the inputs of the B200 FE driver (cmad_b200/fe_driver.py).

Follows `build_fe_problem_from_deck` (cmad/cli/common.py:293-437) for the sections on the path:
``discretization`` (mesh, time schedule, quadrature override), ``residuals.global residual``
(small_disp_equilibrium, displacement or mixed u-p, global Newton settings and line search),
``residuals.local residual`` (small_elastic_plastic, local Newton + line search settings,
per-block materials in the ``{value, active, transform}`` leaf form of
cmad/io/params_builder.py), ``dirichlet bcs.expression`` and ``qoi.name``.  Defaults are the
reference's (cmad/io/deck.py:44-92).  Sections outside the path (output writers, Neumann BCs,
body forces, iterative linear solvers) raise ``NotImplementedError`` when they would change the
result, and are ignored when they only configure output.

Meshes: classic-netCDF Exodus files are read with SciPy (``examples/meshes/notch.exo``);
``.npz`` files hold ``nodes`` / ``conn``.  The unit-cube meshes of ``examples/make_cube_mesh.py``
(``cube_{hex,tet}_{n}.exo``, HDF5-based netCDF-4, not readable here) are regenerated from their
file name.  Side sets are the coordinate-extreme sets ``{x,y,z}{min,max}_sides`` that script
writes and that ``build coordinate sidesets`` adds (cmad/fem/mesh.py:585-636)."""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np
import yaml

from . import fe_driver as drv, fe_mesh
from .material import NewtonSettings, material_from_values
from .parameters import Parameters

_LS_DEFAULT = {"max evals": 4, "sufficient decrease": 1.0e-4, "min backtrack factor": 0.5,
               "max backtrack factor": 0.9}
_GLOBAL_DEFAULT = {"nonlinear max iters": 10, "nonlinear absolute tol": 1.0e-12, "nonlinear relative tol": 1.0e-12}
_LOCAL_DEFAULT = {"nonlinear max iters": 20, "nonlinear absolute tol": 1.0e-12, "nonlinear relative tol": 1.0e-12}
_SAFE = {k: getattr(math, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "pi", "tanh", "sinh", "cosh")}
_SAFE.update({"abs": np.abs, "where": np.where, "minimum": np.minimum, "maximum": np.maximum})
for _k in ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh"):
    _SAFE[_k] = getattr(np, _k)


def load_deck(path: str) -> dict:
    with open(path) as f:
        deck = yaml.safe_load(f)
    if not isinstance(deck, dict):
        raise ValueError(f"deck top-level must be a mapping: {path}")
    if len(deck) == 1 and isinstance(next(iter(deck.values())), dict) and "problem" in next(iter(deck.values())):
        deck = next(iter(deck.values()))                      # Calibr8-style single-key wrapper
    return deck


# ------------------------------------------------------------------------------ parameters
def split_parameters(node):
    """``{value, active?, transform?}`` leaves -> parallel (values, active, transforms) trees;
    bare scalars / lists are inactive and untransformed (cmad/io/params_builder.py)."""
    def coerce(v):
        if isinstance(v, list):
            return np.asarray(v, dtype=np.float64)
        return float(v) if isinstance(v, int) and not isinstance(v, bool) else v

    def transform(spec):
        if spec is None:
            return None
        if isinstance(spec, dict) and "bounds" in spec:
            return np.asarray(spec["bounds"], dtype=np.float64)
        if isinstance(spec, dict) and "log" in spec:
            return np.asarray([spec["log"]], dtype=np.float64)
        raise ValueError(f"unknown transform spec: {spec!r}")

    if isinstance(node, dict) and "value" in node:
        return coerce(node["value"]), bool(node.get("active", False)), transform(node.get("transform"))
    if isinstance(node, dict):
        out = ({}, {}, {})
        for k, v in node.items():
            for tree, part in zip(out, split_parameters(v)):
                tree[k] = part
        return out
    return coerce(node), False, None


def material_parameters(section: dict) -> Parameters:
    """Parameters of one block; ``rotation matrix`` defaults to the identity
    (SmallElasticPlastic.material_defaults, cmad/cli/common.py:68-79)."""
    merged = dict(section)
    merged.setdefault("rotation matrix", np.eye(3).tolist())
    values, active, transforms = split_parameters(merged)
    return Parameters(values, active, transforms)


# ------------------------------------------------------------------------------ mesh
def read_mesh(path: str):
    """``(nodes, {block name: connectivity})`` of a deck's ``mesh file``."""
    m = re.fullmatch(r"cube_(hex|tet)_(\d+)\.exo", os.path.basename(path))
    if path.endswith(".npz") and os.path.exists(path):
        z = np.load(path)
        conn = z["conn"] if "conn" in z.files else z["tets"]
        return np.asarray(z["nodes"], float), {str(z["block"]) if "block" in z.files else "block_1": np.asarray(conn, np.int64)}
    if os.path.exists(path):
        try:
            from scipy.io import netcdf_file
            with netcdf_file(path, "r", mmap=False) as f:
                v = f.variables
                nodes = (np.array(v["coord"][:], dtype=np.float64).T if "coord" in v else
                         np.stack([np.array(v[k][:], dtype=np.float64) for k in ("coordx", "coordy", "coordz")], axis=1))
                blocks = {}
                names = None
                if "eb_names" in v:
                    names = ["".join(c.decode() for c in row if c not in (b"", b"\x00")).strip()
                             for row in np.array(v["eb_names"][:])]
                k = 1
                while f"connect{k}" in v:
                    nm = names[k - 1] if names and names[k - 1] else f"block_{k}"
                    blocks[nm] = np.array(v[f"connect{k}"][:], dtype=np.int64) - 1
                    k += 1
            return nodes, blocks
        except Exception:
            if not m:
                raise
    if m:                                                        # examples/make_cube_mesh.py
        n = int(m.group(2))
        nodes, conn = fe_mesh.structured_hex_mesh((n, n, n))
        if m.group(1) == "tet":
            conn = fe_mesh.split_hex_to_tets(conn)
        return nodes, {"all": conn}
    raise FileNotFoundError(f"mesh file not found or not readable here: {path}")


def coordinate_side_nodes(nodes: np.ndarray, name: str) -> np.ndarray:
    """Nodes of the coordinate-extreme side set ``{x,y,z}{min,max}_sides``."""
    m = re.fullmatch(r"([xyz])(min|max)_sides", name)
    if not m:
        raise NotImplementedError(f"side set {name!r}: only coordinate-extreme side sets are resolved here")
    ax = "xyz".index(m.group(1))
    c = nodes[:, ax]
    ext = c.min() if m.group(2) == "min" else c.max()
    tol = 1e-9 * max(float(np.ptp(c)), 1e-300)
    return np.flatnonzero(np.abs(c - ext) <= tol)


def scalar_expression(expr) -> Callable[..., Any]:
    """``f(x, y, z, t)`` from a deck expression string (cmad/io/expressions.py, restated with a
    fixed whitelist of names)."""
    code = compile(str(expr), "<deck expression>", "eval")
    for name in code.co_names:
        if name not in _SAFE and name not in ("x", "y", "z", "t"):
            raise ValueError(f"deck expression {expr!r}: unknown name {name!r}")
    return lambda x, y, z, t: eval(code, {"__builtins__": {}}, {**_SAFE, "x": x, "y": y, "z": z, "t": t})


# ------------------------------------------------------------------------------ the problem
@dataclass
class DeckProblem:
    name: str
    nodes: np.ndarray
    conn: np.ndarray
    block: str
    mixed: bool
    stab_mult: float
    volume_degree: int | None
    parameters: Parameters
    local_newton: NewtonSettings
    nonlinear: dict
    t_schedule: np.ndarray
    bc_entries: list = field(default_factory=list)          # (component, side set, expression)
    qoi: str | None = None

    @property
    def values(self):
        return self.parameters.values

    def material(self):
        return material_from_values(self.values)

    def arrays(self, device="cpu"):
        return fe_mesh.block_arrays(self.nodes, self.conn, device=device, mixed=self.mixed,
                                    volume_degree=self.volume_degree)

    def dirichlet_bcs(self) -> drv.DirichletBCs:
        """Prescribed dofs in ascending order with their value expressions (a dof named by
        several entries must agree - cmad/fem/bcs.py - and is prescribed once)."""
        fns, by_dof = [], {}
        for comp, sideset, expr in self.bc_entries:
            f = scalar_expression(expr)
            for nd in coordinate_side_nodes(self.nodes, sideset):
                by_dof.setdefault(int(nd) * 3 + int(comp), (len(fns), int(nd)))
            fns.append(f)
        idx = np.array(sorted(by_dof), dtype=np.int64)
        which = [by_dof[int(i)] for i in idx]
        X = self.nodes

        def values(t):
            return np.array([float(fns[k](X[nd, 0], X[nd, 1], X[nd, 2], t)) for k, nd in which])
        return drv.DirichletBCs(idx, values)

    def pattern(self, arr):
        ur, uc, scatter = fe_mesh.coo_dedup(arr.elem_eq.cpu().numpy(),
                                            arr.elem_eq_p.cpu().numpy() if self.mixed else None)
        return drv.SparsePattern(ur, uc, arr.n_dofs), scatter


def fe_problem_from_deck(path: str) -> DeckProblem:
    deck = load_deck(path)
    if deck.get("problem", {}).get("type") != "fe":
        raise NotImplementedError("only problem.type: fe decks are built here (material-point decks: cmad_b200.primal)")
    base = os.path.dirname(os.path.abspath(path))
    disc = deck["discretization"]
    nodes, blocks = read_mesh(os.path.join(base, disc["mesh file"]))
    if len(blocks) != 1:
        raise NotImplementedError("multi-block meshes: build one DeckProblem per block")
    block, conn = next(iter(blocks.items()))
    gr = {**_GLOBAL_DEFAULT, **deck["residuals"]["global residual"]}
    if gr.get("type") != "small_disp_equilibrium" or str(gr.get("def_type", "full_3d")).lower() != "full_3d":
        raise NotImplementedError("global residual: small_disp_equilibrium / full_3d only")
    loc = {**_LOCAL_DEFAULT, **deck["residuals"]["local residual"]}
    if loc.get("type") != "small_elastic_plastic":
        raise NotImplementedError(f"local residual type {loc.get('type')!r} is outside the B200 FE path")
    mats = loc["materials"]
    if set(mats) != {block}:
        raise ValueError(f"residuals.local residual.materials keys ({sorted(mats)}) must match mesh element "
                         f"blocks ({[block]})")
    for k in ("surface flux bcs", "body forces"):
        if deck.get(k):
            raise NotImplementedError(f"deck section {k!r} is outside the B200 FE path")
    ls_type = deck.get("linear solver", {}).get("type", "direct")
    if ls_type != "direct":
        raise NotImplementedError("linear solver: the host side of this path solves with a direct factorisation")
    mixed = bool(gr.get("mixed", False))
    quad = (disc.get("quadrature") or {}).get("volume degree")
    if mixed and quad is not None and int(quad) < 2:
        raise ValueError(f"residuals.global residual: mixed requires volume quadrature degree >= 2; got {quad}")
    volume_degree = int(quad) if quad is not None else (2 if mixed else None)      # cli/common.py:379-391
    lls = {**_LS_DEFAULT, **(loc.get("line search") or {})}
    local_newton = NewtonSettings(mode="traced", max_iters=int(loc["nonlinear max iters"]),
                                  abs_tol=float(loc["nonlinear absolute tol"]), rel_tol=float(loc["nonlinear relative tol"]),
                                  ls_max_evals=int(lls["max evals"]), ls_sufficient_decrease=float(lls["sufficient decrease"]),
                                  ls_min_backtrack=float(lls["min backtrack factor"]),
                                  ls_max_backtrack=float(lls["max backtrack factor"]))
    gls = {**_LS_DEFAULT, **{k: v for k, v in (gr.get("line search") or {}).items() if k != "print"}}
    nonlinear = {"max iters": int(gr["nonlinear max iters"]), "abs tol": float(gr["nonlinear absolute tol"]),
                 "rel tol": float(gr["nonlinear relative tol"]), "line search": gls}
    if "times" in disc:
        ts = np.asarray(disc["times"], dtype=np.float64).ravel()
    elif "times file" in disc:
        p = os.path.join(base, disc["times file"])
        ts = np.asarray(np.load(p) if p.endswith(".npy") else np.loadtxt(p), dtype=np.float64).ravel()
    else:
        ts = np.arange(int(disc["num steps"]) + 1, dtype=np.float64) * float(disc["step size"])
    entries = []
    for name, (resid, eq, sideset, expr) in (deck.get("dirichlet bcs", {}).get("expression", {}) or {}).items():
        if resid != "equilibrium":
            raise NotImplementedError(f"dirichlet bcs.expression.{name}: only the equilibrium residual is prescribed here")
        if not 0 <= int(eq) < 3:
            raise ValueError(f"dirichlet bcs.expression.{name}: eq {eq} out of range for residual 'equilibrium'")
        entries.append((int(eq), str(sideset), expr))
    return DeckProblem(name=str(deck["problem"].get("name", "")), nodes=nodes, conn=conn, block=block, mixed=mixed,
                       stab_mult=float(gr.get("stabilization multiplier", 1.0)), volume_degree=volume_degree,
                       parameters=material_parameters(mats[block]), local_newton=local_newton, nonlinear=nonlinear,
                       t_schedule=ts, bc_entries=entries, qoi=(deck.get("qoi") or {}).get("name"))


def run_primal(problem: DeckProblem, device="cuda:0", step_qoi=None):
    """`cmad primal deck.yaml` for an FE deck over the CUDA kernels: returns
    ``(U_steps, xi_last, J, logs)`` of the quasi-static drive (cmad/fem/driver.py:149-253)."""
    import torch
    from . import fe
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("run_primal needs a CUDA device (cmad_b200 has no CPU fallback)")
    arr_h = problem.arrays()
    arr = arr_h.to(dev)
    pattern, scatter = problem.pattern(arr_h)
    k_plan = fe.SegmentPlan(scatter, len(pattern.rows), device=dev)
    mat = problem.material()
    if problem.mixed:
        asm = drv.cuda_assembler_mixed(mat, problem.local_newton, arr, fe.mixed_r_plan(arr, device=dev), k_plan,
                                       problem.stab_mult)
    else:
        r_plan = fe.SegmentPlan(arr_h.elem_eq.numpy().reshape(-1), arr.n_dofs, device=dev)
        asm = drv.cuda_assembler(mat, problem.local_newton, arr, r_plan, k_plan)
    bcs = problem.dirichlet_bcs()
    xi0 = torch.zeros((arr.n_elems, arr.n_ip, 7), dtype=torch.float64, device=dev)
    if step_qoi is None and problem.qoi == "fe_displacement_l2" and not problem.mixed:
        wdet = (arr_h.det * arr_h.quad_w[None, :]).numpy()
        vol, T = wdet.sum(), float(problem.t_schedule[-1] - problem.t_schedule[0])
        N, eq = arr_h.N.numpy(), arr_h.elem_eq.numpy()
        step_qoi = lambda U, t, tp: (t - tp) / (T * vol) * drv.displacement_l2_step(N, wdet, eq, U)   # noqa: E731
    return drv.fe_quasistatic_drive(drv.DeviceEmbeddedBCs(asm, pattern, bcs, dev), pattern, bcs,
                                    np.zeros(arr.n_dofs), xi0, problem.t_schedule, problem.nonlinear, step_qoi)
