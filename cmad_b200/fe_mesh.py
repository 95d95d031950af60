"""Synthetic FE workloads for the element-block kernels: structured hex / tet
meshes and the mesh-derived arrays the kernels consume, in the REFERENCE's
layouts.  Host-side set-up that runs once per mesh (torch used as an array
library, on whatever device is asked for); nothing here is on the hot path.

Layouts produced (reference file:line):
  * connectivity / node numbering of ``StructuredHexMesh`` (cmad/fem/mesh.py:424-514):
    node id = C-order index of the (nx+1, ny+1, nz+1) grid, element
    ``e = i*ny*nz + j*nz + k``, hex_linear node order (cmad/fem/interpolants.py:15-24);
  * geometry cache (cmad/fem/precompute.py:170-271): ``iso_jac_det (n_e, n_ip)``,
    ``grad_N_phys (n_e, n_ip, n_b, 3)``, shared ``quad_w (n_ip)``, ``N (n_ip, n_b)``;
  * default rules hex degree 2 (2x2x2 Gauss-Legendre, cmad/fem/quadrature.py:70-93) and
    tet degree 1 (centroid, w = 1/6, :128-129) per cmad/fem/fe_problem.py:35-38;
  * equation indices ``eq = offset + node*3 + comp`` in (basis, comp) order
    (cmad/fem/assembly.py:142-165) = ``u_gather_eq`` = ``r_scatter_eq``
    (cmad/fem/kernel_arrays.py:68-80);
  * COO pattern and dedup scatter (cmad/fem/assembly.py:970-1070).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

_HEX_NODE_XI = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                         [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def structured_hex_mesh(divisions, lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
    """``(nodes (n_nodes,3), connectivity (n_elems,8))`` of a Cartesian box."""
    nx, ny, nz = (int(d) for d in divisions)
    if min(nx, ny, nz) < 1:
        raise ValueError(f"divisions must all be >= 1; got {divisions}")
    axes = [np.linspace(o, o + L, n + 1) for o, L, n in zip(origin, lengths, (nx, ny, nz))]
    nodes = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    vid = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int64).reshape(nx + 1, ny + 1, nz + 1)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.stack([vid[i + a, j + b, k + c] for a, b, c in corners], axis=-1).reshape(-1, 8)
    return nodes, conn


# six tets around the body diagonal 0-6 of a hex, every one positively oriented
# on a positively oriented hex (volume checked in tests)
_HEX_TO_TET = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6],
                        [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]], dtype=np.int64)


def split_hex_to_tets(conn_hex: np.ndarray) -> np.ndarray:
    """Each hex -> 6 tets (tet ``6*e + t``), cf. ``hex_to_tet_split`` (cmad/fem/mesh.py:516-580)."""
    return conn_hex[:, _HEX_TO_TET].reshape(-1, 4)


def hex_quadrature_deg2():
    g = 1.0 / np.sqrt(3.0)
    x1 = np.array([-g, g])
    xi = np.stack(np.meshgrid(x1, x1, x1, indexing="ij"), axis=-1).reshape(-1, 3)
    return xi, np.ones(8)


def tet_quadrature_deg1():
    return np.array([[0.25, 0.25, 0.25]]), np.array([1.0 / 6.0])


def hex_quadrature(degree: int):
    """Tensor Gauss-Legendre rule on [-1, 1]^3 exact to ``degree`` per axis: ceil((d+1)/2)
    points per axis, first axis slowest (cmad/fem/quadrature.py:70-93)."""
    n = (int(degree) + 2) // 2
    x1, w1 = np.polynomial.legendre.leggauss(n)
    xi = np.stack(np.meshgrid(x1, x1, x1, indexing="ij"), axis=-1).reshape(-1, 3)
    w = (w1[:, None, None] * w1[None, :, None] * w1[None, None, :]).reshape(-1)
    return xi, w


def tet_quadrature(degree: int):
    """Rules on the unit tetrahedron (weights sum to 1/6): degree 1 = centroid; degree 2 = the
    symmetric 4-point rule (the rule the reference's mixed formulation needs on tets,
    cmad/cli/common.py:379-391; cmad/fem/quadrature.py:251-272)."""
    if degree == 1:
        return tet_quadrature_deg1()
    if degree == 2:
        a, b = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0, (5.0 - np.sqrt(5.0)) / 20.0
        xi = np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]])
        return xi, np.full(4, 1.0 / 24.0)
    raise ValueError(f"tet_quadrature: degree {degree} is not tabulated here (1, 2)")


def hex_linear_shapes(xi: np.ndarray):
    """``N (n_ip, 8)`` and reference gradients ``(n_ip, 8, 3)`` of the trilinear hex."""
    t = 1.0 + xi[:, None, :] * _HEX_NODE_XI[None, :, :]          # (n_ip, 8, 3)
    N = t.prod(axis=2) / 8.0
    g = np.stack([_HEX_NODE_XI[None, :, 0] * t[:, :, 1] * t[:, :, 2],
                  _HEX_NODE_XI[None, :, 1] * t[:, :, 0] * t[:, :, 2],
                  _HEX_NODE_XI[None, :, 2] * t[:, :, 0] * t[:, :, 1]], axis=2) / 8.0
    return N, g


def tet_linear_shapes(xi: np.ndarray):
    N = np.stack([1.0 - xi.sum(axis=1), xi[:, 0], xi[:, 1], xi[:, 2]], axis=1)
    g = np.broadcast_to(np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]]),
                        (xi.shape[0], 4, 3)).copy()
    return N, g


@dataclass
class FEBlockArrays:
    """Mesh-derived arrays of one element block (≙ the block's slices of the
    reference's ``FEKernelArrays`` + ``BlockIPGeometryCache``).  The mixed u-p
    formulation adds the pressure block's equation indices (block-major dofs:
    ``eq_p = 3 n_nodes + node``, cmad/fem/dof.py) and the element sizes."""
    elem_eq: torch.Tensor    # (n_e, n_b*3) int32
    grad_N: torch.Tensor     # (n_e, n_ip, n_b, 3) float64, physical frame
    det: torch.Tensor        # (n_e, n_ip) float64
    quad_w: torch.Tensor     # (n_ip,) float64
    N: torch.Tensor          # (n_ip, n_b) float64 (shared)
    n_dofs: int
    elem_eq_p: torch.Tensor | None = None   # (n_e, n_b) int32   (mixed u-p only)
    h: torch.Tensor | None = None           # (n_e,) float64 RMS edge length (mixed u-p only)

    @property
    def n_elems(self) -> int:
        return int(self.elem_eq.shape[0])

    @property
    def n_basis(self) -> int:
        return int(self.grad_N.shape[2])

    @property
    def n_ip(self) -> int:
        return int(self.grad_N.shape[1])

    @property
    def mixed(self) -> bool:
        return self.elem_eq_p is not None

    def to(self, device) -> "FEBlockArrays":
        mv = lambda t: None if t is None else t.to(device)  # noqa: E731
        return FEBlockArrays(self.elem_eq.to(device), self.grad_N.to(device), self.det.to(device),
                             self.quad_w.to(device), self.N.to(device), self.n_dofs,
                             mv(self.elem_eq_p), mv(self.h))

    def slice(self, lo: int, hi: int) -> "FEBlockArrays":
        """Contiguous element range (multi-GPU partition by element index)."""
        sl = lambda t: None if t is None else t[lo:hi].contiguous()  # noqa: E731
        return FEBlockArrays(self.elem_eq[lo:hi].contiguous(), self.grad_N[lo:hi].contiguous(),
                             self.det[lo:hi].contiguous(), self.quad_w, self.N, self.n_dofs,
                             sl(self.elem_eq_p), sl(self.h))


# local edges (cmad/fem/mesh.py _HEX_LOCAL_EDGES / _TET_LOCAL_EDGES; only the set matters here)
_HEX_EDGES = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4],
                       [0, 4], [1, 5], [2, 6], [3, 7]])
_TET_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])


def element_rms_edge_sizes(nodes, conn) -> np.ndarray:
    """``h[e] = sqrt(mean_k l_k^2)`` over the element's edges (cmad/fem/mesh.py:624-636):
    the length scale of the mixed formulation's pressure stabilisation."""
    conn = np.asarray(conn)
    edges = _HEX_EDGES if conn.shape[1] == 8 else _TET_EDGES
    X = np.asarray(nodes, dtype=np.float64)[conn[:, edges]]          # (n_e, n_edges, 2, 3)
    return np.sqrt(np.mean(np.sum((X[:, :, 1] - X[:, :, 0]) ** 2, axis=-1), axis=-1))


def block_arrays(nodes, conn, device="cpu", chunk: int = 1 << 20, mixed: bool = False,
                 volume_degree: int | None = None) -> FEBlockArrays:
    """Geometry cache + equation indices of one block (tet4 if ``conn`` has 4
    columns, hex8 if 8), as ``precompute_block_geometry`` builds them.  ``mixed``:
    the u-p formulation (one pressure dof per node after the 3 n_nodes displacement dofs).
    ``volume_degree``: the deck's ``discretization.quadrature.volume degree`` override (None =
    the family defaults hex 2 / tet 1, cmad/fem/fe_problem.py:35-38)."""
    conn = np.asarray(conn)
    n_b = conn.shape[1]
    if n_b == 8:
        xi, w = hex_quadrature_deg2() if volume_degree is None else hex_quadrature(volume_degree)
        N, gref = hex_linear_shapes(xi)
    elif n_b == 4:
        xi, w = tet_quadrature_deg1() if volume_degree is None else tet_quadrature(volume_degree)
        N, gref = tet_linear_shapes(xi)
    else:
        raise ValueError(f"unsupported element with {n_b} nodes")
    dev = torch.device(device)
    X_all = torch.as_tensor(np.asarray(nodes, dtype=np.float64))
    gref_t = torch.as_tensor(gref).to(dev)
    conn_t = torch.as_tensor(conn.astype(np.int64))
    n_e = conn.shape[0]
    grad_N = torch.empty((n_e, len(w), n_b, 3), dtype=torch.float64, device=dev)
    det = torch.empty((n_e, len(w)), dtype=torch.float64, device=dev)
    for lo in range(0, n_e, chunk):
        hi = min(n_e, lo + chunk)
        X = X_all[conn_t[lo:hi]].to(dev)                               # (c, n_b, 3)
        jac = torch.einsum("eai,paj->epij", X, gref_t)                 # iso_jac[e,p,i,j]
        det[lo:hi] = torch.linalg.det(jac)
        grad_N[lo:hi] = torch.einsum("pnj,epji->epni", gref_t, torch.linalg.inv(jac))
    eq = (conn_t[:, :, None] * 3 + torch.arange(3)[None, None, :]).reshape(n_e, n_b * 3)
    n_nodes = int(np.asarray(nodes).shape[0])
    arr = FEBlockArrays(eq.to(torch.int32).to(dev), grad_N, det, torch.as_tensor(w).to(dev),
                        torch.as_tensor(N).to(dev), n_nodes * 3)
    if mixed:
        arr.n_dofs = n_nodes * 4
        arr.elem_eq_p = (conn_t + 3 * n_nodes).to(torch.int32).to(dev).contiguous()
        arr.h = torch.as_tensor(element_rms_edge_sizes(nodes, conn)).to(dev)
    return arr


def coo_pattern(elem_eq: np.ndarray):
    """With-duplicates ``(rows, cols)`` in ``assemble_element_block``'s emit order
    ``(elem, row dof, col dof)`` (cmad/fem/assembly.py:970-1023)."""
    eq = np.asarray(elem_eq, dtype=np.int64)
    n_e, nd = eq.shape
    rows = np.broadcast_to(eq[:, :, None], (n_e, nd, nd)).reshape(-1)
    cols = np.broadcast_to(eq[:, None, :], (n_e, nd, nd)).reshape(-1)
    return rows, cols


def coo_pattern_mixed(elem_eq: np.ndarray, elem_eq_p: np.ndarray):
    """With-duplicates ``(rows, cols)`` of the mixed u-p block in the emit order of
    ``assemble_element_block``: the (u,u), (u,p), (p,u), (p,p) streams one after the other,
    each ``(elem, row dof, col dof)`` (cmad/fem/assembly.py:722-732, 970-1023)."""
    eqs = [np.asarray(elem_eq, dtype=np.int64), np.asarray(elem_eq_p, dtype=np.int64)]
    rows, cols = [], []
    for r in range(2):
        for s_ in range(2):
            n_e, nr = eqs[r].shape
            nc = eqs[s_].shape[1]
            rows.append(np.broadcast_to(eqs[r][:, :, None], (n_e, nr, nc)).reshape(-1))
            cols.append(np.broadcast_to(eqs[s_][:, None, :], (n_e, nr, nc)).reshape(-1))
    return np.concatenate(rows), np.concatenate(cols)


def coo_dedup(elem_eq: np.ndarray, elem_eq_p: np.ndarray | None = None):
    """``(unique_rows, unique_cols, coo_dedup_scatter)`` as ``assembled_coo_dedup``
    (cmad/fem/assembly.py:1026-1070): lex-sorted unique pattern + the map from each
    with-duplicates triplet to its slot."""
    rows, cols = coo_pattern(elem_eq) if elem_eq_p is None else coo_pattern_mixed(elem_eq, elem_eq_p)
    n = int(max(rows.max(), cols.max())) + 1
    key = rows * n + cols
    uniq, inverse = np.unique(key, return_inverse=True)
    return (uniq // n).astype(np.int64), (uniq % n).astype(np.int64), inverse.astype(np.int64)


def synthetic_displacement(nodes: np.ndarray, t: float, seed: int = 42, ramp: float = 0.003,
                           noise: float = 1e-4) -> np.ndarray:
    """Uniaxial ramp ``u_x = ramp * t * x`` plus a random nodal perturbation
    (seed/perturbation as tests/fem/test_assembly_coupled.py:177-178)."""
    rng = np.random.default_rng(seed)
    U = noise * rng.standard_normal(nodes.shape)
    U[:, 0] += ramp * t * nodes[:, 0]
    return U.reshape(-1)
