"""Line-by-line CPU restatement (torch fp64 + torch.func AD) of CMAD's
per-integration-point constitutive update.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Derivatives here are
AD-derived exactly where the reference uses ``jax.jacfwd / jacrev / grad``, so
this file is an *independent* check of the hand-derived CUDA kernels.

All ``file:line`` citations are relative to the reference tree
(``/root/reference``).  Parity status: pinned on KA1..KA7;
parity against actual JAX output is unpinned (JAX not installable here).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np
import torch
from torch.func import grad, jacfwd, jacrev

DT = torch.float64

# --------------------------------------------------------------------------
# enums (cmad/models/deformation_types.py:4-9, cmad/models/deriv_types.py)
# --------------------------------------------------------------------------
FULL_3D, PLANE_STRAIN, PLANE_STRESS, UNIAXIAL_STRESS, PURE_SHEAR = range(5)


def def_type_ndims(def_type: int) -> int:
    """cmad/models/deformation_types.py:12-20."""
    if def_type == FULL_3D:
        return 3
    if def_type in (PLANE_STRAIN, PLANE_STRESS):
        return 2
    if def_type in (UNIAXIAL_STRESS, PURE_SHEAR):
        return 1
    raise NotImplementedError


# --------------------------------------------------------------------------
# parameter pytrees (cmad/parameters/parameters.py)
# --------------------------------------------------------------------------
def flatten_tree(tree: Any, path: tuple = ()) -> list[tuple[tuple, Any]]:
    """JAX pytree flatten of nested dicts: keys in *sorted* order, leaves in
    depth-first order (``jax.tree_util`` dict semantics used by
    ``ravel_pytree`` in parameters.py:214-227).  ``None`` is a leaf here (the
    reference flattens transforms with ``is_leaf=lambda x: x is None``)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree.keys()):
            out += flatten_tree(tree[k], path + (k,))
        return out
    return [(path, tree)]


def tree_map(fn: Callable, tree: Any, *rest: Any) -> Any:
    if isinstance(tree, dict):
        return {k: tree_map(fn, tree[k], *[r[k] for r in rest]) for k in tree}
    return fn(tree, *rest)


def leaf_size(v: Any) -> int:
    return int(np.size(np.asarray(v)))


def ravel_params(values: dict) -> np.ndarray:
    """``ravel_pytree(values)[0]`` (parameters.py:214)."""
    return np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1)
                           for _, v in flatten_tree(values)])


def unravel_params(values_like: dict, flat: np.ndarray) -> dict:
    """Inverse of :func:`ravel_params` (``reconstruct_from_flat``)."""
    it = [0]

    def rebuild(t):
        if isinstance(t, dict):
            return {k: rebuild(t[k]) for k in sorted(t.keys())}
        n = leaf_size(t)
        seg = flat[it[0]:it[0] + n]
        it[0] += n
        a = np.asarray(t)
        return seg.reshape(a.shape) if a.ndim else seg[0]

    return rebuild(values_like)


def bounds_transform(v, bounds):
    """parameters.py:27-42 (canonical -> native)."""
    span = 0.5 * (bounds[1] - bounds[0])
    mean = 0.5 * (bounds[0] + bounds[1])
    return span * v + mean


def log_transform(v, ref):
    """parameters.py:45-54 (canonical -> native)."""
    return ref[0] * math.exp(v) if not torch.is_tensor(v) else ref[0] * torch.exp(v)


def transform_from_canonical(v, active, transform):
    """parameters.py:141-152."""
    if active and transform is not None:
        if len(transform) == 2:
            return bounds_transform(v, transform)
        if len(transform) == 1:
            return log_transform(v, transform)
        raise ValueError
    return v


def first_deriv_transform(value, transform):
    """parameters.py:94-102."""
    if transform is None:
        return 1.0
    if len(transform) == 2:
        return 0.5 * (transform[1] - transform[0])
    if len(transform) == 1:
        return value
    raise ValueError


class OracleParameters:
    """Restatement of ``Parameters`` (parameters.py:177-401): values / active
    flags / transforms pytrees, flat order, active-column selection and the
    canonical<->native chain rule."""

    def __init__(self, values: dict, active_flags: dict | None = None,
                 transforms: dict | None = None):
        self.values = values
        self._active_flags = active_flags
        self._transforms = transforms
        leaves = flatten_tree(values)
        self.names = ["/".join(p) for p, _ in leaves]
        self.flat_param_sizes = [leaf_size(v) for _, v in leaves]
        self.num_params = int(sum(self.flat_param_sizes))
        if active_flags is not None:
            flags, trs = [], []
            fl = flatten_tree(active_flags)
            tl = flatten_tree(transforms)
            for (_, v), (_, a), (_, t) in zip(leaves, fl, tl):
                n = leaf_size(v)
                flags += [bool(a)] * n          # parameters.py:67-87
                trs += [t] * n
            self._flat_active_flags = np.array(flags)
            self.active_idx = np.arange(self.num_params)[self._flat_active_flags]
            self.num_active_params = int(self._flat_active_flags.sum())
            self._flat_transforms = trs
            self._flat_active_transforms = [trs[i] for i in self.active_idx]
        else:
            self.active_idx = np.zeros(0, dtype=int)
            self.num_active_params = 0

    def flat_values(self) -> np.ndarray:
        return ravel_params(self.values)

    def flat_active_values(self, return_canonical: bool = False) -> np.ndarray:
        """parameters.py:304-317."""
        flat = self.flat_values()
        if not return_canonical:
            return flat[self.active_idx]
        out = []
        for i in self.active_idx:
            v, t = flat[i], self._flat_transforms[i]
            if t is None:
                out.append(v)
            elif len(t) == 2:
                span = 0.5 * (t[1] - t[0]); mean = 0.5 * (t[0] + t[1])
                out.append(min(1.0, max(-1.0, (v - mean) / span)))
            else:
                out.append(math.log(v / t[0]))
        return np.array(out)

    def set_active_values_from_flat(self, flat_active, are_canonical=True):
        """parameters.py:278-301."""
        flat = self.flat_values()
        flat[self.active_idx] = flat_active
        vals = unravel_params(self.values, flat)
        if are_canonical:
            vals = tree_map(
                lambda v, a, t: (transform_from_canonical(float(v), a, t)
                                 if np.ndim(v) == 0 else v),
                vals, self._active_flags, self._transforms)
        self.values = vals

    def transform_grad(self, g: np.ndarray) -> None:
        """parameters.py:326-331 (in place)."""
        av = self.flat_values()[self.active_idx]
        for i in range(self.num_active_params):
            g[i] = first_deriv_transform(av[i], self._flat_active_transforms[i]) * g[i]

    def active_params_jacobian(self, jac_tree: dict, num_eqns: int) -> np.ndarray:
        """parameters.py:368-377: reshape each leaf block to (num_eqns, -1),
        hstack in flatten order, keep active columns."""
        cols = [np.asarray(v).reshape(num_eqns, -1) for _, v in flatten_tree(jac_tree)]
        return np.hstack(cols)[:, self.active_idx]


def to_torch_tree(values: dict) -> dict:
    return tree_map(lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64)), values)


# --------------------------------------------------------------------------
# tensor <-> vector packing (cmad/models/var_types.py:43-55, 73-84)
# --------------------------------------------------------------------------
def sym_tensor_from_vector(v):
    return torch.stack([torch.stack([v[0], v[1], v[2]]),
                        torch.stack([v[1], v[3], v[4]]),
                        torch.stack([v[2], v[4], v[5]])])


def vector_from_sym_tensor(T):
    return torch.stack([T[0, 0], T[0, 1], T[0, 2], T[1, 1], T[1, 2], T[2, 2]])


# --------------------------------------------------------------------------
# elasticity (cmad/models/elastic_constants.py:9-18, 54-104;
#             cmad/models/elastic_stress.py:14-21, 24-40, 71-72)
# --------------------------------------------------------------------------
def lame_from_params(el: dict):
    keys = frozenset(k for k in ("E", "nu", "mu", "kappa", "lambda") if k in el)
    if len(keys) != 2:
        raise ValueError(f"need exactly two elastic constants; got {sorted(keys)}")
    if keys == {"lambda", "mu"}:
        return el["lambda"], el["mu"]
    if keys == {"E", "nu"}:
        E, nu = el["E"], el["nu"]
        return E * nu / ((1. + nu) * (1. - 2. * nu)), E / (2. * (1. + nu))
    if keys == {"mu", "kappa"}:
        mu, ka = el["mu"], el["kappa"]
        return ka - 2. * mu / 3., mu
    if keys == {"E", "mu"}:
        E, mu = el["E"], el["mu"]
        return mu * (E - 2. * mu) / (3. * mu - E), mu
    if keys == {"E", "kappa"}:
        E, ka = el["E"], el["kappa"]
        return 3. * ka * (3. * ka - E) / (9. * ka - E), 3. * ka * E / (9. * ka - E)
    if keys == {"mu", "nu"}:
        mu, nu = el["mu"], el["nu"]
        return 2. * mu * nu / (1. - 2. * nu), mu
    if keys == {"kappa", "nu"}:
        ka, nu = el["kappa"], el["nu"]
        return 3. * ka * nu / (1. + nu), 3. * ka * (1. - 2. * nu) / (2. * (1. + nu))
    if keys == {"lambda", "nu"}:
        la, nu = el["lambda"], el["nu"]
        return la, la * (1. - 2. * nu) / (2. * nu)
    if keys == {"lambda", "kappa"}:
        la, ka = el["lambda"], el["kappa"]
        return la, 3. * (ka - la) / 2.
    if keys == {"E", "lambda"}:
        E, la = el["E"], el["lambda"]
        R = (E ** 2 + 9. * la ** 2 + 2. * E * la) ** 0.5
        return la, (E - 3. * la + R) / 4.
    raise ValueError


def isotropic_linear_elastic_stress(ee, params):
    """elastic_stress.py:14-21 (form used by the elastic-plastic models)."""
    lam, mu = lame_from_params(params["elastic"])
    return lam * torch.trace(ee) * torch.eye(3, dtype=DT) + 2. * mu * ee


def isotropic_linear_elastic_cauchy_stress(F, params):
    """elastic_stress.py:24-40 (kappa/dev form used by ``Elastic``)."""
    I = torch.eye(3, dtype=DT)
    grad_u = F - I
    eps = 0.5 * (grad_u + grad_u.T)
    tr = torch.trace(eps)
    dev = eps - tr / 3. * I
    lam, mu = lame_from_params(params["elastic"])
    kappa = lam + 2. * mu / 3.
    return kappa * tr * I + 2. * mu * dev


def two_mu_scale_factor(params):
    """elastic_stress.py:71-72."""
    return 2. * lame_from_params(params["elastic"])[1]


# --------------------------------------------------------------------------
# effective stresses (cmad/models/effective_stress.py:30-37, 40-52, 168-177)
# --------------------------------------------------------------------------
def J2_effective_stress(cauchy, plastic_params=None):
    hydro = torch.trace(cauchy) / 3.
    s = cauchy - hydro * torch.eye(3, dtype=DT)
    snorm = torch.sqrt(torch.sum(s * s))
    return math.sqrt(3. / 2.) * snorm


def hill_effective_stress(cauchy, plastic_params):
    h = plastic_params["effective stress"]["hill"]
    F, G, H, L, M, N = h["F"], h["G"], h["H"], h["L"], h["M"], h["N"]
    return torch.sqrt(F * (cauchy[1, 1] - cauchy[2, 2]) ** 2
                      + G * (cauchy[2, 2] - cauchy[0, 0]) ** 2
                      + H * (cauchy[0, 0] - cauchy[1, 1]) ** 2
                      + L * (cauchy[2, 1] ** 2 + cauchy[1, 2] ** 2)
                      + M * (cauchy[2, 0] ** 2 + cauchy[0, 2] ** 2)
                      + N * (cauchy[1, 0] ** 2 + cauchy[0, 1] ** 2))


def hosford_effective_stress(cauchy, plastic_params):
    """effective_stress.py:167-177: diagonal entries only, scaled by the von
    Mises stress before the power."""
    vm = J2_effective_stress(cauchy, plastic_params)
    a = plastic_params["effective stress"]["hosford"]["a"]
    sc = cauchy / vm
    d01 = torch.abs(sc[0, 0] - sc[1, 1]) ** a
    d12 = torch.abs(sc[1, 1] - sc[2, 2]) ** a
    d20 = torch.abs(sc[2, 2] - sc[0, 0]) ** a
    return vm * (0.5 * (d01 + d12 + d20)) ** (a ** -1)


BARLAT_TENSOR_KEYS = ("12", "13", "21", "23", "31", "32", "44", "55", "66")


def barlat_effective_stress(cauchy, plastic_params):
    """Yld2004-18p, effective_stress.py:55-84 -> verification/functions.py:71-154: two linear
    images of the stress (`sp_*`, `dp_*` coefficient sets; the 3x3 normal blocks annihilate the
    hydrostatic part, the shear entries are scaled entry by entry), their eigenvalues through
    `eigh` of the symmetrised image, phi = (1/4 sum_ij |S'_i - S''_j|^a)^(1/a)."""
    b = plastic_params["effective stress"]["barlat"]

    def image(pre):
        c12, c13, c21, c23, c31, c32, c44, c55, c66 = (b[f"{pre}_{k}"] for k in BARLAT_TENSOR_KEYS)
        d = torch.stack([cauchy[0, 0], cauchy[1, 1], cauchy[2, 2]])
        ul = torch.stack([torch.stack([c12 + c13, -2. * c12 + c13, c12 - 2. * c13]),
                          torch.stack([-2. * c21 + c23, c21 + c23, c21 - 2. * c23]),
                          torch.stack([-2. * c31 + c32, c31 - 2. * c32, c31 + c32])]) / 3.
        n = ul @ d
        return torch.stack([torch.stack([n[0], c44 * cauchy[0, 1], c66 * cauchy[0, 2]]),
                            torch.stack([c44 * cauchy[1, 0], n[1], c55 * cauchy[1, 2]]),
                            torch.stack([c66 * cauchy[2, 0], c55 * cauchy[2, 1], n[2]])])

    def eigvals(S):
        return torch.linalg.eigh(0.5 * (S + S.T))[0]          # jnp.linalg.eigh symmetrises its input

    e1, e2 = eigvals(image("sp")), eigvals(image("dp"))
    a = b["a"]
    return (0.25 * torch.sum(torch.abs(e1[:, None] - e2[None, :]) ** a)) ** (1. / a)


def effective_stress_fun(kind: str):
    """effective_stress.py:16-27."""
    return {"J2": J2_effective_stress, "hill": hill_effective_stress,
            "hosford": hosford_effective_stress, "barlat": barlat_effective_stress}[kind]


# --------------------------------------------------------------------------
# hardening (cmad/models/hardening.py:9-34)
# --------------------------------------------------------------------------
def combined_hardening(alpha, hardening_params: dict):
    total = 0.
    for htype in hardening_params:  # dict order; the sum is order-insensitive to 1 ulp
        hp = hardening_params[htype]
        if htype == "voce":
            total = total + hp["S"] * (1. - torch.exp(-hp["D"] * alpha))
        elif htype == "linear":
            total = total + hp["K"] * alpha
        else:
            raise NotImplementedError(htype)
    return total


# --------------------------------------------------------------------------
# models: flat residual / cauchy (state flattened in block order, i.e. the
# ``ravel_pytree`` order used by make_newton_solve, nonlinear_solver.py:103)
# --------------------------------------------------------------------------
@dataclass
class ModelSpec:
    """Static description of a reference ``Model`` instance."""
    kind: str = "small_elastic_plastic"      # or "elastic"
    def_type: int = FULL_3D
    yield_tol: float = 1e-14                 # small_elastic_plastic.py:116
    uniaxial_stress_idx: int = 0
    effective_stress: str | None = None      # None -> first key of the params subtree

    def block_sizes(self) -> list[int]:
        if self.kind in ("small_elastic_plastic", "small_rate_elastic_plastic"):   # small_elastic_plastic.py:126-180
            b = [6, 1]
        elif self.kind == "elastic":                   # elastic.py:57-97
            b = [6]
        else:
            raise NotImplementedError(self.kind)
        if self.def_type == PLANE_STRESS:
            b.append(1)
        elif self.def_type == UNIAXIAL_STRESS:
            b.append(2)
            if self.kind == "small_rate_elastic_plastic":      # + off-axis delta strains, small_rate_elastic_plastic.py:189-199
                b.append(3)
        elif self.def_type != FULL_3D:
            raise NotImplementedError
        return b

    @property
    def num_dofs(self) -> int:
        return sum(self.block_sizes())

    def init_xi(self) -> np.ndarray:
        """small_elastic_plastic.py:139-180 / elastic.py:76-97: zeros for the
        tensor/scalar state, ones for the stretches."""
        b = self.block_sizes()
        x = np.zeros(sum(b))
        n0 = 6 if self.kind == "elastic" else 7
        x[n0:] = 1.0
        if self.kind == "small_rate_elastic_plastic" and self.def_type == UNIAXIAL_STRESS:
            x[n0 + 2:] = 0.0                                   # the off-axis delta strains start at zero
        return x


def _gather_F(x, grad_u, spec: ModelSpec, local_off: int):
    """cmad/models/kinematics.py:10-53.  ``x[local_off:]`` holds the stretch
    block (``local_var_idx`` in the reference)."""
    if spec.def_type == FULL_3D:
        return torch.eye(3, dtype=DT) + grad_u
    if spec.def_type == PLANE_STRESS:
        F2 = torch.eye(2, dtype=DT) + grad_u
        F = torch.zeros(3, 3, dtype=DT)
        top = torch.cat([F2, torch.zeros(2, 1, dtype=DT)], dim=1)
        bot = torch.cat([torch.zeros(1, 2, dtype=DT), x[local_off:local_off + 1].reshape(1, 1)], dim=1)
        return torch.cat([top, bot], dim=0) + 0. * F
    if spec.def_type == UNIAXIAL_STRESS:
        F11 = (torch.eye(1, dtype=DT) + grad_u)[0, 0]
        st = x[local_off:local_off + 2]
        idx = spec.uniaxial_stress_idx
        if idx == 0:
            d = torch.stack([F11, st[0], st[1]])
        elif idx == 1:
            d = torch.stack([st[0], F11, st[1]])
        else:
            d = torch.stack([st[0], st[1], F11])
        return torch.diag(d)
    raise NotImplementedError


def _sep_elastic_strain(x, params, grad_u, spec: ModelSpec):
    """small_elastic_plastic.py:33-64."""
    F = _gather_F(x, grad_u, spec, 7)
    ep = sym_tensor_from_vector(x[0:6])
    gu = F - torch.eye(3, dtype=DT)
    eps = 0.5 * (gu + gu.T)
    Q = params["rotation matrix"]
    if spec.def_type == UNIAXIAL_STRESS:
        g = Q @ ep @ Q.T
        c = torch.stack([torch.stack([eps[0, 0], g[0, 1], g[0, 2]]),
                         torch.stack([g[1, 0], eps[1, 1], g[1, 2]]),
                         torch.stack([g[2, 0], g[2, 1], eps[2, 2]])])
        em = Q.T @ c @ Q
    else:
        em = Q.T @ eps @ Q
    return em - ep


def _sep_es_kind(params, spec):
    if spec.effective_stress is not None:
        return spec.effective_stress
    return next(iter(params["plastic"]["effective stress"]))   # small_elastic_plastic.py:193-198


def sep_yield_fun_and_normal(x, params, grad_u, spec: ModelSpec):
    """small_elastic_plastic.py:67-92."""
    pl = params["plastic"]
    Y = pl["flow stress"]["initial yield"]["Y"]
    ee = _sep_elastic_strain(x, params, grad_u, spec)
    cauchy = isotropic_linear_elastic_stress(ee, params)
    phi_fun = effective_stress_fun(_sep_es_kind(params, spec))
    phi = phi_fun(cauchy, pl)
    alpha = x[6]
    sigma_flow = Y + combined_hardening(alpha, pl["flow stress"]["hardening"])
    f = (phi - sigma_flow) / two_mu_scale_factor(params)
    n = grad(phi_fun)(cauchy, pl)
    return cauchy, f, n


def sep_is_plastic(f, tol: float):
    """cmad/models/paths.py:26."""
    return torch.logical_or(f > tol, torch.abs(f) < tol)


def sep_residual(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """small_elastic_plastic.py:238-302 (+ paths.py:26-27)."""
    ep = sym_tensor_from_vector(x[0:6])
    ep_prev = sym_tensor_from_vector(x_prev[0:6])
    alpha, alpha_prev = x[6], x_prev[6]
    dgamma = alpha - alpha_prev
    sig_m, f, n = sep_yield_fun_and_normal(x, params, grad_u, spec)
    Ce_t = ep - ep_prev
    Ce = torch.cat([vector_from_sym_tensor(Ce_t), dgamma.reshape(1)])
    Cp_t = Ce_t - dgamma * n
    Cp = torch.cat([vector_from_sym_tensor(Cp_t), f.reshape(1)])
    if spec.def_type in (PLANE_STRESS, UNIAXIAL_STRESS):
        sc = two_mu_scale_factor(params)
        Q = params["rotation matrix"]
        g = Q @ sig_m @ Q.T
        if spec.def_type == PLANE_STRESS:
            Cs = (g[2, 2] / sc).reshape(1)
        else:
            i0, i1 = [i for i in range(3) if i != spec.uniaxial_stress_idx]
            Cs = torch.stack([g[i0, i0], g[i1, i1]]) / sc
        Ce = torch.cat([Ce, Cs])
        Cp = torch.cat([Cp, Cs])
    return torch.where(sep_is_plastic(f, spec.yield_tol), Cp, Ce)


def sep_cauchy(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """small_elastic_plastic.py:308-321."""
    ee = _sep_elastic_strain(x, params, grad_u, spec)
    sig_m = isotropic_linear_elastic_stress(ee, params)
    Q = params["rotation matrix"]
    return Q @ sig_m @ Q.T


def elastic_residual(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """cmad/models/elastic.py:139-173."""
    cauchy = sym_tensor_from_vector(x[0:6])
    F = _gather_F(x, grad_u, spec, 6)
    sc = two_mu_scale_factor(params)
    C = vector_from_sym_tensor(cauchy - isotropic_linear_elastic_cauchy_stress(F, params)) / sc
    if spec.def_type == PLANE_STRESS:
        C = torch.cat([C, (cauchy[2, 2] / sc).reshape(1)])
    elif spec.def_type == UNIAXIAL_STRESS:
        C = torch.cat([C, torch.stack([cauchy[1, 1], cauchy[2, 2]]) / sc])
    return C


def elastic_cauchy(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """cmad/models/elastic.py:188-195."""
    return sym_tensor_from_vector(x[0:6])


def rate_residual(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """small_rate_elastic_plastic.py:250-346 (FULL_3D): state = [cauchy(6), alpha]; the trial
    stress increment comes from eps(U) - eps(U_prev) (:34-77), the yield function and its
    normal are evaluated at the state's own stress (:80-100)."""
    cauchy = sym_tensor_from_vector(x[0:6])
    cauchy_prev = sym_tensor_from_vector(x_prev[0:6])
    alpha, alpha_prev = x[6], x_prev[6]
    Q = params["rotation matrix"]
    I3 = torch.eye(3, dtype=DT)
    if spec.def_type == FULL_3D:
        gu, gup = grad_u, grad_u_prev
    else:                                                      # :41-51 (the def-types' stretches enter F)
        gu = _gather_F(x, grad_u, spec, 7) - I3
        gup = _gather_F(x_prev, grad_u_prev, spec, 7) - I3
    de = 0.5 * (gu + gu.T) - 0.5 * (gup + gup.T)
    if spec.def_type == UNIAXIAL_STRESS:                       # :57-70: off-diagonals = the fourth state block
        d = x[9:12]
        de = torch.stack([torch.stack([de[0, 0], d[0], d[1]]), torch.stack([d[0], de[1, 1], d[2]]),
                          torch.stack([d[1], d[2], de[2, 2]])])
    trial = isotropic_linear_elastic_stress(Q.T @ de @ Q, params)
    dgamma = alpha - alpha_prev
    sc = two_mu_scale_factor(params)
    pl = params["plastic"]
    phi_fun = effective_stress_fun(_sep_es_kind(params, spec))
    f = (phi_fun(cauchy, pl) - (pl["flow stress"]["initial yield"]["Y"]
                                + combined_hardening(alpha, pl["flow stress"]["hardening"]))) / sc
    n = grad(phi_fun)(cauchy, pl)
    Ce = torch.cat([vector_from_sym_tensor(cauchy - cauchy_prev - trial) / sc, dgamma.reshape(1)])
    dc = trial - isotropic_linear_elastic_stress(dgamma * n, params)
    Cp = torch.cat([vector_from_sym_tensor(cauchy - cauchy_prev - dc) / sc, f.reshape(1)])
    if spec.def_type != FULL_3D:                               # :296-345: constraints on the GLOBAL stress increment
        gt, gd = Q @ trial @ Q.T, Q @ dc @ Q.T
        if spec.def_type == PLANE_STRESS:
            Ce = torch.cat([Ce, (gt[2, 2] / sc).reshape(1)]); Cp = torch.cat([Cp, (gd[2, 2] / sc).reshape(1)])
        else:
            assert spec.uniaxial_stress_idx == 0
            rows = lambda g: torch.stack([g[1, 1], g[2, 2], g[0, 1], g[0, 2], g[1, 2]]) / sc  # noqa: E731
            Ce = torch.cat([Ce, rows(gt)]); Cp = torch.cat([Cp, rows(gd)])
    return torch.where(sep_is_plastic(f, spec.yield_tol), Cp, Ce)


def rate_cauchy(x, x_prev, params, grad_u, grad_u_prev, spec: ModelSpec):
    """small_rate_elastic_plastic.py:351-359."""
    Q = params["rotation matrix"]
    return Q @ sym_tensor_from_vector(x[0:6]) @ Q.T


def residual_fun(spec: ModelSpec):
    return {"small_elastic_plastic": sep_residual, "elastic": elastic_residual,
            "small_rate_elastic_plastic": rate_residual}[spec.kind]


def cauchy_fun(spec: ModelSpec):
    return {"small_elastic_plastic": sep_cauchy, "elastic": elastic_cauchy,
            "small_rate_elastic_plastic": rate_cauchy}[spec.kind]


# --------------------------------------------------------------------------
# line search (cmad/util/line_search.py:40-46, 74-85, 95-189)
# --------------------------------------------------------------------------
DEFAULT_LINE_SEARCH_SETTINGS = {
    "max evals": 4, "sufficient decrease": 1.0e-4,
    "min backtrack factor": 0.5, "max backtrack factor": 0.9,
}


def quad_min(phi_0, dphi_0, a, phi):
    """line_search.py:74-85."""
    denom = 2.0 * (phi - phi_0 - dphi_0 * a)
    if denom == 0.0:
        return 0.5 * a
    return -dphi_0 * a * a / denom


def line_search(eval_fn, phi_0, dphi_0, settings, init_aux):
    """line_search.py:95-189, quadratic branch (``slope is None``, the local
    Newton's case).  ``eval_fn(alpha) -> (phi, aux)``.  Returns
    ``(alpha, aux, n_evals)``."""
    max_evals = settings["max evals"]
    c1 = settings["sufficient decrease"]
    bmin = settings["min backtrack factor"]
    bmax = settings["max backtrack factor"]
    armijo_slope = c1 * dphi_0
    n, alpha, accepted, aux = 0, 1.0, False, init_aux
    best_alpha, best_phi, best_aux = 1.0, math.inf, init_aux
    while n < max_evals and not accepted:
        phi, aux = eval_fn(alpha)
        finite = math.isfinite(phi)
        if finite and phi < best_phi:
            best_alpha, best_phi, best_aux = alpha, phi, aux
        accepted = finite and (phi <= phi_0 + alpha * armijo_slope)
        alpha_model = quad_min(phi_0, dphi_0, alpha, phi)
        # jnp.clip(x, lo, hi) == min(max(x, lo), hi); NaN propagates
        lo, hi = bmin * alpha, bmax * alpha
        if alpha_model != alpha_model:
            alpha_contracted = alpha_model
        else:
            alpha_contracted = min(max(alpha_model, lo), hi)
        if not accepted:
            alpha = alpha_contracted if finite else 0.5 * alpha
        n += 1
    if accepted:
        return alpha, aux, n
    return best_alpha, best_aux, n


# --------------------------------------------------------------------------
# local Newton solvers (cmad/models/nonlinear_solver.py)
# --------------------------------------------------------------------------
@dataclass
class NewtonInfo:
    iters: int = 0
    converged: bool = False
    C_norm: float = 0.0
    flag_entry: int = 0     # is_plastic at x0      (paths.py:26)
    flag_exit: int = 0      # is_plastic at x*      (paths.py:26)
    ls_evals: int = 0


def _flag(x, params, grad_u, spec) -> int:
    if spec.kind != "small_elastic_plastic":
        return 0
    _, f, _ = sep_yield_fun_and_normal(x, params, grad_u, spec)
    return int(bool(sep_is_plastic(f, spec.yield_tol)))


def newton_traced(x_prev, params, grad_u, grad_u_prev, spec: ModelSpec,
                  max_iters: int = 10, abs_tol: float = 1e-14,
                  rel_tol: float = 1e-14,
                  line_search_settings: dict | None = None):
    """``make_newton_solve`` primal (nonlinear_solver.py:102-155): the FE
    path's and the ``jvp`` MP strategy's local Newton.  Returns
    ``(x, NewtonInfo)`` - the reference returns only ``x``."""
    ls = {**DEFAULT_LINE_SEARCH_SETTINGS, **(line_search_settings or {})}
    res = residual_fun(spec)
    x_prev = torch.as_tensor(x_prev, dtype=DT)

    def rflat(x):
        return res(x, x_prev, params, grad_u, grad_u_prev, spec)

    info = NewtonInfo(flag_entry=_flag(x_prev, params, grad_u, spec))
    x = x_prev.clone()
    C = rflat(x)                                           # :111
    n0 = float(torch.linalg.norm(C))                       # :112
    ii, converged = 0, False
    with np.errstate(invalid="ignore", divide="ignore"):
        while ii < max_iters and not converged:            # :135-137
            n = float(torch.linalg.norm(C))                # :142
            rel = np.float64(n) / np.float64(n0)           # :143 (0/0 -> NaN)
            if rel < rel_tol or n < abs_tol:               # :149
                converged = True
                continue
            J = jacfwd(rflat)(x)                           # :122
            delta = torch.linalg.solve(J, C)               # :123

            def eval_fn(a, x=x, delta=delta):
                Ct = rflat(x - a * delta)
                return float(0.5 * (Ct @ Ct)), Ct          # :125-127

            CC = float(C @ C)
            a, C, ne = line_search(eval_fn, 0.5 * CC, -CC, ls, C)   # :129-131
            info.ls_evals += ne
            x = x - a * delta
            ii += 1                                        # :132
    info.iters, info.converged = ii, converged
    info.C_norm = float(torch.linalg.norm(C))
    info.flag_exit = _flag(x, params, grad_u, spec)
    return x, info


def newton_imperative(x_prev, params, grad_u, grad_u_prev, spec: ModelSpec,
                      max_iters: int = 10, abs_tol: float = 1e-14,
                      rel_tol: float = 1e-14, x_init=None):
    """Imperative ``newton_solve(model)`` (nonlinear_solver.py:14-85) with the
    default ``max_ls_evals=0`` (no line search): the MP primal / objective /
    adjoint / direct paths.  Starts from the model's current ``xi`` which, after
    ``advance_xi``/``set_xi_to_init_vals``, equals ``xi_prev``."""
    res = residual_fun(spec)
    x_prev = torch.as_tensor(x_prev, dtype=DT)
    x = x_prev.clone() if x_init is None else torch.as_tensor(x_init, dtype=DT).clone()

    def rflat(x):
        return res(x, x_prev, params, grad_u, grad_u_prev, spec)

    info = NewtonInfo(flag_entry=_flag(x, params, grad_u, spec))
    ii, converged, n0, n = 0, False, 1.0, 0.0
    while ii < max_iters and not converged:
        C = rflat(x)
        n = float(np.linalg.norm(C.numpy()))               # :34
        if ii == 0:
            n0, rel = n, 1.0                               # :36-38
        else:
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.float64(n) / np.float64(n0)       # :40
        if rel < rel_tol or n < abs_tol:                   # :42-44
            converged = True
            break
        J = jacfwd(rflat)(x)
        delta = np.linalg.solve(J.numpy(), -C.numpy())     # :52
        x = x + torch.as_tensor(delta)                     # :53
        ii += 1
    info.iters, info.converged, info.C_norm = ii, converged, n
    info.flag_exit = _flag(x, params, grad_u, spec)
    return x, info


# --------------------------------------------------------------------------
# AD products of the model (cmad/models/model.py:121-166, 316-374)
# --------------------------------------------------------------------------
def dC_dxi(x, x_prev, params, grad_u, grad_u_prev, spec):
    return jacfwd(residual_fun(spec), argnums=0)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dC_dxi_prev(x, x_prev, params, grad_u, grad_u_prev, spec):
    return jacfwd(residual_fun(spec), argnums=1)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dC_dparams(x, x_prev, params, grad_u, grad_u_prev, spec) -> dict:
    """model.py:128 (jacrev): pytree parallel to params, leaves (n_xi, *leaf)."""
    return jacrev(residual_fun(spec), argnums=2)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dC_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec):
    """model.py:129: the ``grad_fields['u']`` leaf of dC/dU, shape (n_xi, nd, nd)."""
    return jacfwd(residual_fun(spec), argnums=3)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dcauchy_dxi(x, x_prev, params, grad_u, grad_u_prev, spec):
    return jacfwd(cauchy_fun(spec), argnums=0)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dcauchy_dparams(x, x_prev, params, grad_u, grad_u_prev, spec) -> dict:
    return jacrev(cauchy_fun(spec), argnums=2)(x, x_prev, params, grad_u, grad_u_prev, spec)


def dcauchy_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec):
    return jacfwd(cauchy_fun(spec), argnums=3)(x, x_prev, params, grad_u, grad_u_prev, spec)


def ift_dxi_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec):
    """nonlinear_solver.py:158-171 specialised to a ``grad_u`` tangent:
    ``dxi/dgrad_u = -A^{-1} dC/dgrad_u`` (n_xi, nd, nd)."""
    A = dC_dxi(x, x_prev, params, grad_u, grad_u_prev, spec)
    B = dC_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec)
    return -torch.linalg.solve(A, B.reshape(A.shape[0], -1)).reshape(B.shape)


def consistent_tangent(x, x_prev, params, grad_u, grad_u_prev, spec):
    """d cauchy / d grad_u |_total  (SURVEY A.5): shape (3,3,nd,nd)."""
    dsx = dcauchy_dxi(x, x_prev, params, grad_u, grad_u_prev, spec)          # (3,3,n)
    dsu = dcauchy_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec)      # (3,3,nd,nd)
    dxu = ift_dxi_dgrad_u(x, x_prev, params, grad_u, grad_u_prev, spec)      # (n,nd,nd)
    return dsu + torch.einsum("ijn,nkl->ijkl", dsx, dxu)


# --------------------------------------------------------------------------
# QoI (cmad/qois/calibration.py:56-66)
# --------------------------------------------------------------------------
def calibration_qoi(x, x_prev, params, grad_u, grad_u_prev, spec, data, weight):
    cauchy = cauchy_fun(spec)(x, x_prev, params, grad_u, grad_u_prev, spec)
    mis = weight * (cauchy - data)
    return 0.5 * torch.sum(mis * mis)


def uniaxial_calibration_qoi(x, x_prev, params, grad_u, grad_u_prev, spec, data, weight):
    """cmad/qois/uniaxial_calibration.py:69-85 with ``uniaxial_stress_idx = spec.uniaxial_stress_idx``
    and ``stretch_var_idx = 2`` (the off-axis stretches at ``x[7:9]``); data / weight (3,) of the step."""
    cauchy = cauchy_fun(spec)(x, x_prev, params, grad_u, grad_u_prev, spec)
    i = spec.uniaxial_stress_idx
    pred = torch.stack([cauchy[i, i], x[7] - 1.0, x[8] - 1.0])
    mis = (pred - data) * weight
    return 0.5 * torch.sum(mis * mis)


def _qoi_of(weight):
    """Calibration for a (3, 3) weight, UniaxialCalibration for per-step weights (3, N+1):
    returns (qoi function, weight-at-step accessor)."""
    w = torch.as_tensor(np.asarray(weight), dtype=DT)
    if tuple(w.shape) == (3, 3):
        return calibration_qoi, (lambda step: w)
    return uniaxial_calibration_qoi, (lambda step: w[:, step])


# --------------------------------------------------------------------------
# MP drivers (cmad/cli/primal.py:129-176, cmad/objectives/mp_objective.py)
# --------------------------------------------------------------------------
def _grad_u_from_F(F_step, spec):
    nd = def_type_ndims(spec.def_type)
    return torch.as_tensor(np.asarray(F_step)[:nd, :nd] - np.eye(nd), dtype=DT)


def mp_primal(parameters: OracleParameters, F: np.ndarray, spec: ModelSpec,
              **newton_kwargs):
    """``run_primal_pass`` (cli/primal.py:129-176): returns xi history
    (N+1, n_xi), cauchy (3,3,N+1), per-step iteration counts, final norms."""
    params = to_torch_tree(parameters.values)
    N = F.shape[-1] - 1
    xi = np.zeros((N + 1, spec.num_dofs)); xi[0] = spec.init_xi()
    cauchy = np.zeros((3, 3, N + 1)); iters = np.zeros(N + 1, dtype=int)
    norms = np.zeros(N + 1); flags = np.zeros(N + 1, dtype=int)
    x = torch.as_tensor(xi[0])
    for step in range(1, N + 1):
        gu = _grad_u_from_F(F[:, :, step], spec)
        gup = _grad_u_from_F(F[:, :, step - 1], spec)
        x_new, info = newton_imperative(x, params, gu, gup, spec, **newton_kwargs)
        cauchy[:, :, step] = cauchy_fun(spec)(x_new, x_new, params, gu, gup, spec).numpy()
        xi[step] = x_new.numpy(); iters[step] = info.iters
        norms[step] = info.C_norm; flags[step] = info.flag_exit
        x = x_new
    return xi, cauchy, iters, norms, flags


def _step_derivs(x, x_prev, params, gu, gup, spec, parameters, data_s, weight, qoi_fn=calibration_qoi):
    A = dC_dxi(x, x_prev, params, gu, gup, spec).numpy()
    B = dC_dxi_prev(x, x_prev, params, gu, gup, spec).numpy()
    dCdp = parameters.active_params_jacobian(
        tree_map(lambda t: t.numpy(), dC_dparams(x, x_prev, params, gu, gup, spec)),
        spec.num_dofs)
    q = lambda *a: qoi_fn(*a, spec, data_s, weight)
    dJdx = jacfwd(q, argnums=0)(x, x_prev, params, gu, gup).numpy().reshape(1, -1)
    dJdp = parameters.active_params_jacobian(
        tree_map(lambda t: t.numpy(), jacrev(q, argnums=2)(x, x_prev, params, gu, gup)), 1)
    return A, B, dCdp, dJdx, dJdp


def mp_objective_adjoint(parameters: OracleParameters, F, data, weight, spec: ModelSpec,
                         flat_active_values=None, are_canonical=True):
    """``MPAdjointObjective`` (mp_objective.py:53-147): J and dJ/dp (canonical
    coordinates after ``transform_grad``)."""
    if flat_active_values is not None:
        parameters.set_active_values_from_flat(np.asarray(flat_active_values, float), are_canonical)
    params = to_torch_tree(parameters.values)
    N = F.shape[-1] - 1
    qoi_fn, w_at = _qoi_of(weight)
    xs = [torch.as_tensor(spec.init_xi())]
    J = 0.0
    for step in range(1, N + 1):                                   # :73-87
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        x, _ = newton_imperative(xs[-1], params, gu, gup, spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        J += float(qoi_fn(x, xs[-1], params, gu, gup, spec, d, w_at(step)))
        xs.append(x)
    Pa = parameters.num_active_params
    g = np.zeros((1, Pa)); hist = np.zeros((spec.num_dofs, 1))
    for step in range(N, 0, -1):                                   # :112-142
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        A, B, dCdp, dJdx, dJdp = _step_derivs(xs[step], xs[step - 1], params, gu, gup, spec,
                                              parameters, d, w_at(step), qoi_fn)
        phi = np.linalg.solve(A.T, -dJdx.T + hist)
        hist = -B.T @ phi
        g += phi.T @ dCdp + dJdp
    g = g.squeeze(0).copy()
    parameters.transform_grad(g)
    return J, g


def mp_objective_direct(parameters: OracleParameters, F, data, weight, spec: ModelSpec,
                        flat_active_values=None, are_canonical=True):
    """``MPDirectObjective`` (mp_objective.py:158-215)."""
    if flat_active_values is not None:
        parameters.set_active_values_from_flat(np.asarray(flat_active_values, float), are_canonical)
    params = to_torch_tree(parameters.values)
    N = F.shape[-1] - 1
    qoi_fn, w_at = _qoi_of(weight)
    Pa = parameters.num_active_params
    x_prev = torch.as_tensor(spec.init_xi())
    J = 0.0; g = np.zeros((1, Pa)); dxdp = np.zeros((spec.num_dofs, Pa))
    for step in range(1, N + 1):
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        x, _ = newton_imperative(x_prev, params, gu, gup, spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        J += float(qoi_fn(x, x_prev, params, gu, gup, spec, d, w_at(step)))
        A, B, dCdp, dJdx, dJdp = _step_derivs(x, x_prev, params, gu, gup, spec, parameters, d, w_at(step), qoi_fn)
        dxdp = np.linalg.solve(A, -dCdp - B @ dxdp)
        g += dJdx @ dxdp + dJdp
        x_prev = x
    g = g.squeeze(0).copy()
    parameters.transform_grad(g)
    return J, g


def second_deriv_transform(value, transform):
    """parameters.py:105-112."""
    if transform is None or len(transform) == 2:
        return 0.0
    if len(transform) == 1:
        return value
    raise ValueError


def transform_hessian(parameters: OracleParameters, H: np.ndarray, grad_native: np.ndarray) -> None:
    """parameters.py:334-357 with diagonal_/off_diagonal_hessian_transform (:115-138), in place."""
    av = parameters.flat_values()[parameters.active_idx]
    tr = parameters._flat_active_transforms
    for i in range(parameters.num_active_params):
        for j in range(parameters.num_active_params):
            if i == j:
                H[i, i] = H[i, i] * first_deriv_transform(av[i], tr[i]) ** 2 \
                    + grad_native[i] * second_deriv_transform(av[i], tr[i])
            elif i < j:
                H[i, j] = H[i, j] * first_deriv_transform(av[i], tr[i]) * first_deriv_transform(av[j], tr[j])
            else:
                H[i, j] = H[j, i]


def _tree_with_active(values: dict, active_idx: np.ndarray, pa: torch.Tensor) -> dict:
    """The params pytree with its ACTIVE flat entries replaced by the entries of the torch
    vector ``pa`` (so AD w.r.t. ``pa`` is AD w.r.t. the active parameters, native values)."""
    pos = {int(i): k for k, i in enumerate(active_idx)}
    it = [0]

    def rebuild(t):
        if isinstance(t, dict):
            return {k: rebuild(t[k]) for k in sorted(t.keys())}
        a = np.asarray(t, dtype=np.float64)
        n = a.size
        ent = [pa[pos[it[0] + q]] if (it[0] + q) in pos else torch.as_tensor(float(a.reshape(-1)[q]), dtype=DT)
               for q in range(n)]
        it[0] += n
        return torch.stack(ent).reshape(a.shape) if a.ndim else ent[0]

    return rebuild(values)


def mp_objective_direct_adjoint(parameters: OracleParameters, F, data, weight, spec: ModelSpec,
                                flat_active_values=None, are_canonical=True,
                                reference_qoi_cross_terms: bool = False):
    """``MPDirectAdjointObjective`` (mp_objective.py:218-343): J, gradient and Hessian in
    canonical coordinates.  Forward pass with storage, the adjoint pass keeping phi_t, then the
    forward direct-adjoint pass.  The thirteen einsum terms of :317-336 are
    ``Z^T (d2(J_t + phi_t . C_t)/dz2) Z`` with ``z = (p, xi_t, xi_{t-1})`` and
    ``Z = [I; dxi_t/dp; dxi_{t-1}/dp]``; the second-derivative tensors come from
    ``torch.func.hessian`` where the reference uses ``jax.hessian / jacrev(jacfwd)``
    (model.py:134-148, qoi.py:47-58).

    ``reference_qoi_cross_terms``: the reference builds the QoI's mixed block
    ``d2J/dxi dparams`` as ``jacrev(jacfwd(qoi_fun, argnums=DXI_PREV), DPARAMS)`` (qoi.py:53-55) -
    differentiating w.r.t. xi_PREV, which the Calibration QoI does not depend on - so that block
    is identically zero there and the terms ``d2J_dp_dxi . dxi_dp`` (:320, :322) drop out.  That
    is exact only while no active parameter enters the Cauchy stress (flow-stress parameters, the
    reference's own tests); with elastic parameters active the reference's Hessian misses them
    (central differences of its own gradient show it: tests/test_hessian.py).  False (default):
    the mathematically complete Hessian; True: the reference's output, entry for entry."""
    from torch.func import hessian
    if flat_active_values is not None:
        parameters.set_active_values_from_flat(np.asarray(flat_active_values, float), are_canonical)
    params = to_torch_tree(parameters.values)
    N = F.shape[-1] - 1
    w = torch.as_tensor(weight, dtype=DT)
    nx, Pa = spec.num_dofs, parameters.num_active_params
    xs = [torch.as_tensor(spec.init_xi())]
    J = 0.0
    for step in range(1, N + 1):
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        x, _ = newton_imperative(xs[-1], params, gu, gup, spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        J += float(calibration_qoi(x, xs[-1], params, gu, gup, spec, d, w))
        xs.append(x)
    g = np.zeros((1, Pa)); hist = np.zeros((nx, 1)); phis = [None] * (N + 1)
    for step in range(N, 0, -1):                                   # :239-268
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        A, B, dCdp, dJdx, dJdp = _step_derivs(xs[step], xs[step - 1], params, gu, gup, spec, parameters, d, w)
        phi = np.linalg.solve(A.T, -dJdx.T + hist)
        phis[step] = phi[:, 0].copy()
        hist = -B.T @ phi
        g += phi.T @ dCdp + dJdp
    g = g.squeeze(0).copy()
    g_native = g.copy()
    parameters.transform_grad(g)
    H = np.zeros((Pa, Pa)); X_prev = np.zeros((nx, Pa))
    pa0 = torch.as_tensor(parameters.flat_values()[parameters.active_idx], dtype=DT)
    res = residual_fun(spec)
    for step in range(1, N + 1):                                   # :275-338
        gu = _grad_u_from_F(F[:, :, step], spec); gup = _grad_u_from_F(F[:, :, step - 1], spec)
        d = torch.as_tensor(data[..., step], dtype=DT)
        A, B, dCdp, dJdx, dJdp = _step_derivs(xs[step], xs[step - 1], params, gu, gup, spec, parameters, d, w)
        X = np.linalg.solve(A, -dCdp - B @ X_prev)
        phi_t = torch.as_tensor(phis[step], dtype=DT)

        def qoi_z(z):
            p = _tree_with_active(parameters.values, parameters.active_idx, z[:Pa])
            return calibration_qoi(z[Pa:Pa + nx], z[Pa + nx:], p, gu, gup, spec, d, w)

        def constraint_z(z):
            p = _tree_with_active(parameters.values, parameters.active_idx, z[:Pa])
            return phi_t @ res(z[Pa:Pa + nx], z[Pa + nx:], p, gu, gup, spec)
        z0 = torch.cat([pa0, xs[step], xs[step - 1]])
        HJ = hessian(qoi_z)(z0).numpy()
        if reference_qoi_cross_terms:                              # qoi.py:53-55, see the docstring
            HJ[:Pa, Pa:Pa + nx] = 0.0
            HJ[Pa:Pa + nx, :Pa] = 0.0
        Hz = HJ + hessian(constraint_z)(z0).numpy()
        Z = np.vstack([np.eye(Pa), X, X_prev])
        H += Z.T @ Hz @ Z
        X_prev = X
    transform_hessian(parameters, H, g_native)
    return J, g, H


# --------------------------------------------------------------------------
# FE per-IP / per-element / per-block (global_residuals/*, fem/assembly.py)
# --------------------------------------------------------------------------
def interpolate_grad_u(U_e, grad_N):
    """global_residuals/interpolation.py:54: ``grad_u[k,j] = sum_a U[a,k] gradN[a,j]``."""
    return U_e.T @ grad_N


def coupled_ip(params, U_e, U_e_prev, x_prev, grad_N, w, dv, spec: ModelSpec,
               newton_settings: dict | None = None, want_tangent: bool = True):
    """``R_and_dR_dU_and_xi`` for the displacement formulation
    (global_residual.py:361-395 + small_disp_equilibrium.py:112-118):
    returns ``R (n_b,3)``, ``dR_dU (n_b,3,n_b,3)``, ``xi``, NewtonInfo."""
    settings = newton_settings or {"abs_tol": 1e-12, "rel_tol": 1e-12, "max_iters": 20}
    U_e = torch.as_tensor(U_e, dtype=DT); U_e_prev = torch.as_tensor(U_e_prev, dtype=DT)
    grad_N = torch.as_tensor(grad_N, dtype=DT)
    gu = interpolate_grad_u(U_e, grad_N); gup = interpolate_grad_u(U_e_prev, grad_N)
    x, info = newton_traced(x_prev, params, gu, gup, spec, **settings)
    xp = torch.as_tensor(x_prev, dtype=DT)
    sigma = cauchy_fun(spec)(x, xp, params, gu, gup, spec)
    R = (grad_N @ sigma) * w * dv
    if not want_tangent:
        return R, None, x, info
    D = consistent_tangent(x, xp, params, gu, gup, spec)           # (3,3,3,3) [j,i,k,l]
    # R[a,i] = gradN[a,j] sigma[j,i];  grad_u[k,l] = U[b,k] gradN[b,l]
    dR = torch.einsum("aj,jikl,bl->aibk", grad_N, D, grad_N) * w * dv
    return R, dR, x, info


def coupled_ip_mixed(params, U_e, p_e, U_e_prev, x_prev, grad_N, N, w, dv, h, spec: ModelSpec,
                     stab_mult: float = 1.0, newton_settings: dict | None = None):
    """``R_and_dR_dU_and_xi`` of the MIXED u-p formulation at one integration point
    (small_disp_equilibrium.py:87-111) for a model whose ``dev_cauchy`` / ``hydro_cauchy`` read the
    LOCAL STATE (SmallRateElasticPlastic, small_rate_elastic_plastic.py:361-376):
    ``sigma = dev(cauchy(xi)) - p I``, ``R_p = (-(p + hydro)/kappa N - tau gradN.grad p) w dv``,
    ``tau = stab h^2 / (2 mu)``.  Total derivatives through the IFT rule: dxi/dU from
    ``ift_dxi_dgrad_u`` (C does not see p).  Returns ``(R_u, R_p, K_uu, K_up, K_pu, K_pp, xi)``."""
    settings = newton_settings or {"abs_tol": 1e-12, "rel_tol": 1e-12, "max_iters": 20}
    U_e = torch.as_tensor(U_e, dtype=DT); U_e_prev = torch.as_tensor(U_e_prev, dtype=DT)
    p_e = torch.as_tensor(p_e, dtype=DT).reshape(-1)
    grad_N = torch.as_tensor(grad_N, dtype=DT); N = torch.as_tensor(N, dtype=DT)
    n_b = U_e.shape[0]
    gu = interpolate_grad_u(U_e, grad_N); gup = interpolate_grad_u(U_e_prev, grad_N)
    x, _ = newton_traced(x_prev, params, gu, gup, spec, **settings)
    xp = torch.as_tensor(x_prev, dtype=DT)
    lam, mu = lame_from_params(params["elastic"])
    kappa = lam + 2. * mu / 3.
    tau = stab_mult * 0.5 * h ** 2 / mu

    def res(U, p, xx):
        g = interpolate_grad_u(U, grad_N)
        cauchy = cauchy_fun(spec)(xx, xp, params, g, gup, spec)
        hydro = torch.trace(cauchy) / 3.
        pr = N @ p
        sigma = cauchy - hydro * torch.eye(3, dtype=DT) - pr * torch.eye(3, dtype=DT)
        Ru = (grad_N @ sigma) * w * dv
        Rp = (-(pr + hydro) / kappa * N - tau * (grad_N @ (grad_N.T @ p))) * w * dv
        return torch.cat([Ru.reshape(-1), Rp])

    R = res(U_e, p_e, x)
    dR_dU = jacfwd(res, argnums=0)(U_e, p_e, x).reshape(4 * n_b, n_b, 3)
    dR_dp = jacfwd(res, argnums=1)(U_e, p_e, x)
    dR_dx = jacfwd(res, argnums=2)(U_e, p_e, x)                                   # (4 n_b, n_xi)
    dx_dgu = ift_dxi_dgrad_u(x, xp, params, gu, gup, spec)                        # (n_xi, 3, 3) [., k, l]
    dx_dU = torch.einsum("xkl,bl->xbk", dx_dgu, grad_N)                           # grad_u[k,l] = U[b,k] gradN[b,l]
    K = dR_dU + torch.einsum("rx,xbk->rbk", dR_dx, dx_dU)
    nu = 3 * n_b
    return (R[:nu].reshape(n_b, 3), R[nu:], K[:nu].reshape(nu, nu), dR_dp[:nu], K[nu:].reshape(n_b, nu),
            dR_dp[nu:], x)


def coupled_element(params, U_e, U_e_prev, xi_prev_ips, grad_N_ips, dets, quad_w,
                    spec: ModelSpec, newton_settings=None, want_tangent=True):
    """``per_element_R_and_K_coupled`` (fem/assembly.py:416-535): scan over IPs,
    ``dv = iso_jac_det`` , ``w = quad_w[ip]``; accumulate R_e, K_e, stack xi."""
    n_ip = len(quad_w)
    n_b = np.asarray(U_e).shape[0]
    R = torch.zeros(n_b, 3, dtype=DT)
    K = torch.zeros(n_b, 3, n_b, 3, dtype=DT) if want_tangent else None
    xs, infos = [], []
    for ip in range(n_ip):
        r, k, x, info = coupled_ip(params, U_e, U_e_prev, xi_prev_ips[ip], grad_N_ips[ip],
                                   float(quad_w[ip]), float(dets[ip]), spec,
                                   newton_settings, want_tangent)
        R = R + r
        if want_tangent:
            K = K + k
        xs.append(x.numpy()); infos.append(info)
    return R, K, np.stack(xs), infos
