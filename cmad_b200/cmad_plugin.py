"""Reference-side adapter: drops the B200 path in behind sandialabs/cmad's own interfaces.

What a CMAD user does (decks, CLI and drivers unchanged):

    import cmad_b200.cmad_plugin as b200
    b200.install()                 # once, before building FE problems; honours CMAD_B200=0

``install()`` hooks the three places the reference exposes for this path:

* ``cmad.fem.assembly.assemble_element_block`` (cmad/fem/assembly.py:616-732): COUPLED blocks
  whose model :func:`supports` go through :func:`assemble_element_block_b200` - same signature,
  same returns ``(R_block, vals, xi_solved)``; everything else falls through to the original.
  The arrays handed to the kernels are the reference's own, unchanged
  (``fe_arrays.u_gather_eq_by_block``, ``geometry_cache`` - cmad/fem/kernel_arrays.py:58-228,
  cmad/fem/precompute.py:58-122): see :func:`block_inputs`.
* ``cmad.fem.fe_problem.build_fe_problem`` (cmad/fem/fe_problem.py:340-462): wrapped only to
  remember each problem's ``local_newton_settings`` (the reference buries them in closures,
  cmad/global_residuals/global_residual.py:292-300).
* the model registry (cmad/io/registry.py:54-103): every supported model name is re-bound to a
  subclass of the SAME class that carries :class:`B200Model` (``b200_material(params)``), so
  ``Model.from_deck`` (cmad/models/model.py:90-108) keeps building what the decks name.

Backends (how the block call reaches the library):

* :class:`FfiBackend` - JAX arrays through the XLA FFI handlers of ``cmad_b200/xla`` (the
  production path; needs JAX, so it is not importable in the build image);
* :class:`TorchBackend` - host / torch arrays through ``ctypes`` -> the C-ABI on a CUDA device
  (what the GPU tests use);
* any object with ``fe_block(inputs, material, newton)`` / ``fe_block_mixed(...)`` - the CPU
  parity test plugs the oracle in here to prove the argument plumbing against the reference.

Nothing in this module computes: it only re-labels arrays.  There is no CPU fallback - without a
backend that reaches the CUDA library the hook raises."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _lib as L
from .material import NewtonSettings, material_from_values

_MODEL_KINDS = {"SmallElasticPlastic": "small_elastic_plastic"}
_LOCAL_NEWTON: dict[int, tuple] = {}      # id(fe_problem) -> (weakref, local_newton_settings given to build_fe_problem)
_installed: dict[str, Any] = {}


class B200Model:
    """Mixin of the registry subclasses: marks a model whose COUPLED blocks the B200 path serves."""
    b200 = True

    def b200_material(self, params) -> L.Material:
        return material_of(self, params)


def _values(params):
    """The parameter ``values`` pytree (params_by_block[block], cmad/fem/assembly.py:39-58) as
    plain Python floats / arrays."""
    if isinstance(params, dict):
        return {k: _values(v) for k, v in params.items()}
    a = np.asarray(params)
    return float(a) if a.ndim == 0 else a.astype(np.float64)


def _kind_of(model):
    return next((_MODEL_KINDS[c.__name__] for c in type(model).__mro__ if c.__name__ in _MODEL_KINDS), None)


def material_of(model, params) -> L.Material:
    kind = _kind_of(model)
    if kind is None:
        raise NotImplementedError(f"{type(model).__name__} is outside the B200 path")
    return material_from_values(_values(params), kind)


def supports(model, gr=None, params=None) -> bool:
    """True when the COUPLED element block of ``model`` (under global residual ``gr``) is served:
    SmallElasticPlastic, FULL_3D, effective stress in {J2, Hill, Hosford, Barlat (K3 / K4 only: no K6 rule)}, hardening within
    {Voce, linear}, ``SmallDispEquilibrium`` displacement or mixed u-p."""
    kind = _kind_of(model)
    if kind is None:
        return False
    dt = getattr(model, "_def_type", getattr(model, "def_type", None))
    if dt is not None and getattr(dt, "name", str(dt)) not in ("FULL_3D", "DefType.FULL_3D"):
        return False
    if gr is not None and type(gr).__name__ != "SmallDispEquilibrium":
        return False
    try:
        p = params if params is not None else model.parameters.values
        material_from_values(_values(p), kind)
    except (NotImplementedError, ValueError, KeyError, AttributeError):
        return False
    return True


def newton_of(fe_problem) -> NewtonSettings:
    """Local Newton settings of the problem's COUPLED blocks: what build_fe_problem was given
    (captured by :func:`install`), else the reference default 20 / 1e-12 / 1e-12 with the default
    line search (cmad/global_residuals/global_residual.py:292-297)."""
    s = getattr(fe_problem, "_b200_local_newton", None)
    if s is None:
        hit = _LOCAL_NEWTON.get(id(fe_problem))
        s = hit[1] if hit is not None and hit[0]() is fe_problem else None     # ids are reused after collection
    s = s or {"abs_tol": 1e-12, "rel_tol": 1e-12, "max_iters": 20}
    return NewtonSettings.from_reference_kwargs(**s)


@dataclass
class BlockInputs:
    """The arrays of one element block exactly as the reference holds them (no copies beyond the
    int32 view of the equation numbers the C-ABI asks for)."""
    elem_eq: Any            # (n_e, 3 n_b) int32   u_gather_eq_by_block[block][0]
    U: Any                  # (n_dofs,)
    xi_prev: Any            # (n_e, n_ip, n_xi)
    grad_N: Any             # (n_e, n_ip, n_b, 3)  geometry_cache[block].per_elem.field_grad_N_phys_per_block[0]
    det: Any                # (n_e, n_ip)          geometry_cache[block].per_elem.iso_jac_det
    quad_w: Any             # (n_ip,)              geometry_cache[block].shared.quad_w
    n_dofs: int
    mixed: bool = False
    elem_eq_p: Any = None   # (n_e, n_b) int32     u_gather_eq_by_block[block][1]
    N: Any = None           # (n_ip, n_b)          geometry_cache[block].shared.field_N_per_block[1]
    h: Any = None           # (n_e,)               geometry_cache[block].per_elem.element_size
    stab_mult: float = 1.0  # gr._stabilization_multiplier


def _i32(a):
    return a.astype("int32") if hasattr(a, "astype") else np.asarray(a, dtype=np.int32)


def block_inputs(fe_problem, fe_arrays, block_name, U_global, xi_prev_per_block) -> BlockInputs:
    geom = fe_arrays.geometry_cache[block_name]
    eqs = fe_arrays.u_gather_eq_by_block[block_name]
    n_e = eqs[0].shape[0]
    gr = fe_problem.gr
    mixed = bool(getattr(gr, "_mixed", False))
    bi = BlockInputs(elem_eq=_i32(eqs[0].reshape(n_e, -1)), U=U_global, xi_prev=xi_prev_per_block,
                     grad_N=geom.per_elem.field_grad_N_phys_per_block[0], det=geom.per_elem.iso_jac_det,
                     quad_w=geom.shared.quad_w, n_dofs=int(fe_problem.dof_map.num_total_dofs), mixed=mixed)
    if mixed:
        bi.elem_eq_p = _i32(eqs[1].reshape(n_e, -1))
        bi.N = geom.shared.field_N_per_block[1]
        bi.h = geom.per_elem.element_size
        bi.stab_mult = float(getattr(gr, "_stabilization_multiplier", 1.0))
    return bi


# ------------------------------------------------------------------------------ backends
class FfiBackend:
    """JAX device arrays -> XLA FFI custom calls (cmad_b200/xla/cmad_b200_xla.cc)."""

    def __init__(self):
        from . import xla
        if not xla.available():
            raise RuntimeError("FfiBackend needs JAX with jax.ffi (not available in this environment)")
        xla.register()
        self._xla = xla

    def fe_block(self, b: BlockInputs, material, newton: NewtonSettings):
        return self._xla.fe_block(b.elem_eq, b.U, b.xi_prev, b.grad_N, b.det, b.quad_w, material, newton.to_struct())

    def fe_block_mixed(self, b: BlockInputs, material, newton: NewtonSettings):
        return self._xla.fe_block_mixed(b.elem_eq, b.elem_eq_p, b.U, b.xi_prev, b.grad_N, b.det, b.quad_w, b.N, b.h,
                                        material, newton.to_struct(), b.stab_mult)


class TorchBackend:
    """Host (NumPy) or torch arrays -> ctypes -> the C-ABI on a CUDA device.  Fails loudly
    without a GPU: there is no CPU path."""

    def __init__(self, device="cuda:0"):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("TorchBackend needs a CUDA device (cmad_b200 has no CPU fallback)")
        self.device = torch.device(device)

    def _arrays(self, b: BlockInputs):
        import torch
        from .fe_mesh import FEBlockArrays
        dev = self.device
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(np.asarray(a)), dtype=dt).to(dev).contiguous()
        arr = FEBlockArrays(t(b.elem_eq, torch.int32), t(b.grad_N, torch.float64), t(b.det, torch.float64),
                            t(b.quad_w, torch.float64),
                            t(b.N, torch.float64) if b.N is not None else None, b.n_dofs)
        if b.mixed:
            arr.elem_eq_p = t(b.elem_eq_p, torch.int32)
            arr.h = t(b.h, torch.float64)
        return arr, t(b.U, torch.float64), t(b.xi_prev, torch.float64)

    def assemble(self, b: BlockInputs, material, newton: NewtonSettings):
        """``(R_block, vals, xi)`` with the library's own deterministic scatter (K5)."""
        from . import fe
        arr, U, xi_prev = self._arrays(b)
        if b.mixed:
            R, vals, xi = fe.assemble_element_block_mixed(material, newton, arr, U, xi_prev, stab_mult=b.stab_mult,
                                                          r_plan=fe.mixed_r_plan(arr, device=self.device))
        else:
            plan = fe.SegmentPlan(np.asarray(b.elem_eq).reshape(-1), b.n_dofs, device=self.device)
            R, vals, xi = fe.assemble_element_block(material, newton, arr, U, xi_prev, r_plan=plan)
        return R.cpu().numpy(), vals.cpu().numpy(), xi.cpu().numpy()


_default_backend = None


def default_backend():
    global _default_backend
    if _default_backend is None:
        from . import xla
        _default_backend = FfiBackend() if xla.available() else TorchBackend()
    return _default_backend


# ------------------------------------------------------------------------------ the hook
def _scatter_add(n, index, values, like):
    if hasattr(like, "at") and not isinstance(like, np.ndarray):        # jax array
        import jax.numpy as jnp
        return jnp.zeros(n).at[jnp.ravel(index)].add(jnp.ravel(values))
    out = np.zeros(n)
    np.add.at(out, np.asarray(index).ravel(), np.asarray(values).ravel())
    return out


def _cat(parts, like):
    if hasattr(like, "at") and not isinstance(like, np.ndarray):
        import jax.numpy as jnp
        return jnp.concatenate([jnp.ravel(p) for p in parts])
    return np.concatenate([np.asarray(p).ravel() for p in parts])


def assemble_element_block_b200(fe_problem, fe_arrays, params_by_block, block_name, U_global, U_prev_global, t,
                                xi_prev_per_block=None, *, backend=None, newton: NewtonSettings | None = None):
    """Same contract as the reference's ``assemble_element_block`` for a COUPLED block
    (cmad/fem/assembly.py:616-732): returns ``(R_block, vals, xi_solved_per_block)``.
    ``U_prev_global`` and ``t`` do not enter SmallElasticPlastic / SmallDispEquilibrium
    without body forces; blocks with forcing functions are not served."""
    if xi_prev_per_block is None:
        raise ValueError(f"COUPLED block '{block_name}' requires xi_prev_per_block; got None")
    if getattr(fe_problem, "forcing_fns_by_block_idx", None):
        raise NotImplementedError("body-force terms are outside the B200 path")
    model = fe_problem.models_by_block[block_name]
    material = material_of(model, params_by_block[block_name])
    nw = newton or newton_of(fe_problem)
    be = backend or default_backend()
    b = block_inputs(fe_problem, fe_arrays, block_name, U_global, xi_prev_per_block)
    if hasattr(be, "assemble"):
        return be.assemble(b, material, nw)
    if not b.mixed:
        R_e, K_e, xi = be.fe_block(b, material, nw)
        R_block = _scatter_add(b.n_dofs, b.elem_eq, R_e, U_global)             # assembly.py:715-720
        return R_block, _cat([K_e], U_global), xi
    R_u, R_p, K_uu, K_up, K_pu, K_pp, xi = be.fe_block_mixed(b, material, nw)
    R_block = _scatter_add(b.n_dofs, b.elem_eq, R_u, U_global) + _scatter_add(b.n_dofs, b.elem_eq_p, R_p, U_global)
    return R_block, _cat([K_uu, K_up, K_pu, K_pp], U_global), xi               # (r, s) emit order, :722-732


def differentiable_block(b: BlockInputs, material, newton: NewtonSettings, active_pid):
    """``f(p_active, U, xi_prev) -> (R_e, K_e, xi)`` of one displacement block as a
    ``jax.custom_vjp``: the rule ``jax.grad`` needs where the reference transposes ``jax.jvp``
    through the FE Newton's IFT rule (cmad/fem/nonlinear_solver.py:490-537).  ``p_active`` (the
    native values of the active parameters, in ``active_pid`` order) only carries the
    cotangent - the primal uses ``material``.  Backward: ``cmadx_fe_block_vjp`` (pbar,
    xibar_prev) + ``cmadx_fe_block_vjp_disp`` (Ubar).  The cotangent of ``K_e`` is ignored (the
    reference never differentiates the tangent on this path).  Needs real JAX."""
    import jax
    import jax.numpy as jnp
    from . import xla
    nw = newton.to_struct()
    pid = np.asarray(active_pid, np.int32)

    @jax.custom_vjp
    def block(p_active, U, xi_prev):
        return xla.fe_block(b.elem_eq, U, xi_prev, b.grad_N, b.det, b.quad_w, material, nw)

    def fwd(p_active, U, xi_prev):
        out = block(p_active, U, xi_prev)
        return out, (U, xi_prev, out[2])

    def bwd(res, cot):
        U, xi_prev, xi_state = res
        Rbar_e, _Kbar, xibar = cot
        eq = jnp.ravel(b.elem_eq)
        # the kernels contract the element rows themselves: hand them the cotangent per global dof
        # of an R that was gathered per element, i.e. Rbar_e scattered back onto the dofs
        Rbar_nodal = jnp.zeros(b.n_dofs).at[eq].add(jnp.ravel(Rbar_e))
        pbar, xibar_prev = xla.fe_block_vjp(b.elem_eq, U, xi_prev, b.grad_N, b.det, b.quad_w, xi_state,
                                            Rbar_nodal, xibar, material, pid)
        Ubar_ip, _ = xla.fe_block_vjp_disp(b.elem_eq, U, xi_prev, b.grad_N, b.det, b.quad_w, xi_state,
                                           Rbar_nodal, xibar, material)
        n_ip = b.det.shape[1]
        rows = jnp.ravel(jnp.repeat(b.elem_eq, n_ip, axis=0))
        Ubar = jnp.zeros(b.n_dofs).at[rows].add(jnp.ravel(Ubar_ip))
        return pbar, Ubar, xibar_prev

    block.defvjp(fwd, bwd)
    return block


# ------------------------------------------------------------------------------ install
def install(backend=None, env: str = "CMAD_B200") -> bool:
    """Hook the reference (must be importable as ``cmad``).  Returns False (and does nothing)
    when the environment variable ``env`` is set to "0"."""
    if os.environ.get(env, "1") == "0":
        return False
    if _installed:
        return True
    import cmad.fem.assembly as assembly
    import cmad.fem.fe_problem as fe_problem_mod
    from cmad.global_residuals.modes import GlobalResidualMode
    from cmad.io import registry

    original = assembly.assemble_element_block
    original_build = fe_problem_mod.build_fe_problem

    def assemble_element_block(fe_problem, fe_arrays, params_by_block, block_name, U_global, U_prev_global, t,
                               xi_prev_per_block=None):
        if fe_problem.modes_by_block[block_name] == GlobalResidualMode.COUPLED and \
                not getattr(fe_problem, "forcing_fns_by_block_idx", None) and \
                supports(fe_problem.models_by_block[block_name], fe_problem.gr, params_by_block[block_name]):
            return assemble_element_block_b200(fe_problem, fe_arrays, params_by_block, block_name, U_global,
                                               U_prev_global, t, xi_prev_per_block, backend=backend)
        return original(fe_problem, fe_arrays, params_by_block, block_name, U_global, U_prev_global, t,
                        xi_prev_per_block)

    def build_fe_problem(*args, **kwargs):
        fp = original_build(*args, **kwargs)
        if kwargs.get("local_newton_settings") is not None:
            settings = dict(kwargs["local_newton_settings"])
            try:
                object.__setattr__(fp, "_b200_local_newton", settings)
            except (AttributeError, TypeError):                  # slotted / frozen problem objects
                import weakref
                _LOCAL_NEWTON[id(fp)] = (weakref.ref(fp), settings)
        return fp

    assembly.assemble_element_block = assemble_element_block
    fe_problem_mod.build_fe_problem = build_fe_problem
    for name in ("small_elastic_plastic",):
        cls = registry.resolve_model(name)
        if not issubclass(cls, B200Model):
            registry._REGISTRY[name] = type("B200" + cls.__name__, (B200Model, cls), {})
    _installed.update(assemble=original, build=original_build, assembly=assembly, fe_problem=fe_problem_mod)
    return True


def uninstall() -> None:
    if not _installed:
        return
    _installed["assembly"].assemble_element_block = _installed["assemble"]
    _installed["fe_problem"].build_fe_problem = _installed["build"]
    _installed.clear()
