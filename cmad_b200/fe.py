"""FE element-block assembly: Python front end of K3/K4 (``cmadx_fe_block_assemble``)
and K5 (``cmadx_segment_sum``).

Mirrors the reference's block-level entry points so a caller of
``cmad.fem.assembly`` can switch over:

  ``assemble_element_block``           cmad/fem/assembly.py:616-732
  ``assemble_element_block_residual``  cmad/fem/assembly.py:735-813
  ``assemble_global`` / ``assemble_global_residual``  cmad/fem/assembly.py:816-968

Arrays are torch CUDA tensors in the reference's layouts; torch only owns the
memory and streams.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib as L
from .comm import all_reduce_sum, resolve as comm_resolve
from .fe_mesh import FEBlockArrays
from .material import NewtonSettings

# for_model's COUPLED defaults (cmad/global_residuals/global_residual.py:292-297)
FE_LOCAL_NEWTON_DEFAULTS = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)


def fe_newton_settings(**kw) -> NewtonSettings:
    """Local-Newton settings of the FE path: traced Newton with line search,
    defaults 20 / 1e-12 / 1e-12 (for_model) unless a deck overrides them
    (cmad/cli/common.py:393-398)."""
    d = dict(FE_LOCAL_NEWTON_DEFAULTS)
    d.update(kw)
    return NewtonSettings(mode="traced", **d)


def closed_form_elastic_material(values: dict) -> L.Material:
    """Material for a CLOSED_FORM block of the ``Elastic`` model with the isotropic linear
    stress (cmad/models/elastic.py, elastic_stress.py:24-42; the mode the deck builder picks for
    ``supports_closed_form_cauchy`` models, e.g. examples/mixed_elastic.yaml): the element
    kernels run their elastic branch only - a J2 surface with an infinite yield stress never
    yields, so ``sigma = Cel eps``, ``D = Cel`` and the local state stays at its zero initial
    value.  Pass ``xi_prev = zeros((n_e, n_ip, 7))``; the returned ``xi`` is zero as well (the
    reference's CLOSED_FORM blocks carry no xi).  Works for the displacement and the mixed u-p
    formulation (``dev_cauchy_closed_form`` / ``hydro_cauchy_closed_form``)."""
    from .material import material_from_values
    v = {"rotation matrix": np.eye(3), "elastic": dict(values["elastic"]),
         "plastic": {"effective stress": {"J2": 0.0},
                     "flow stress": {"initial yield": {"Y": float("inf")}, "hardening": {}}}}
    return material_from_values(v)


class SegmentPlan:
    """Deterministic segment-sum plan (K5): ``out[s] = sum vals[i] for seg[i] == s``
    in increasing ``i``.  Built once per mesh from the reference's scatter maps
    (``r_scatter_eq.ravel()`` or ``coo_dedup_scatter``)."""

    def __init__(self, seg_of_item, n_segments: int, device=None):
        seg = np.ascontiguousarray(
            seg_of_item.cpu().numpy() if isinstance(seg_of_item, torch.Tensor) else seg_of_item,
            dtype=np.int64).reshape(-1)
        self.n_items, self.n_segments = int(seg.size), int(n_segments)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = L.lib().cmadx_segment_plan_create(seg.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   self.n_items, self.n_segments, C.byref(self._h))
        L.check(rc, "cmadx_segment_plan_create")

    def sum(self, vals: torch.Tensor, out: torch.Tensor | None = None, accumulate: bool = False,
            stream: torch.cuda.Stream | None = None) -> torch.Tensor:
        if vals.dtype != torch.float64 or not vals.is_contiguous() or vals.numel() != self.n_items:
            raise ValueError(f"vals: expected contiguous float64 with {self.n_items} entries")
        if vals.device != self.device:
            raise ValueError("vals must live on the plan's device")
        if out is None:
            if accumulate:
                raise ValueError("accumulate=True needs an existing `out`")
            out = torch.empty(self.n_segments, dtype=torch.float64, device=self.device)
        elif out.dtype != torch.float64 or not out.is_contiguous() or out.numel() != self.n_segments:
            raise ValueError(f"out: expected contiguous float64 with {self.n_segments} entries")
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            rc = L.lib().cmadx_segment_sum(self._h, C.c_void_p(vals.data_ptr()), C.c_void_p(out.data_ptr()),
                                           int(bool(accumulate)), C.c_void_p(s.cuda_stream))
        L.check(rc, "cmadx_segment_sum")
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().cmadx_segment_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _fe_struct(arrays: FEBlockArrays, U, xi_prev, out: dict, U_prev=None) -> L.FeBlock:
    b = L.FeBlock()
    b.U_prev = U_prev.data_ptr() if U_prev is not None else None
    b.n_elems, b.n_dofs = arrays.n_elems, arrays.n_dofs
    b.n_basis, b.n_ip = arrays.n_basis, arrays.n_ip
    b.elem_eq, b.U, b.xi_prev = arrays.elem_eq.data_ptr(), U.data_ptr(), xi_prev.data_ptr()
    b.grad_N, b.det, b.quad_w = arrays.grad_N.data_ptr(), arrays.det.data_ptr(), arrays.quad_w.data_ptr()
    for name in ("xi", "R_elem", "K_elem", "R_global", "sigma", "iters", "flags"):
        setattr(b, name, out[name].data_ptr() if out.get(name) is not None else None)
    return b


def fe_block_launch(material: L.Material, newton: NewtonSettings, arrays: FEBlockArrays,
                    U_global: torch.Tensor, xi_prev: torch.Tensor, outputs=("xi", "R_elem", "K_elem"),
                    out: dict | None = None, stream: torch.cuda.Stream | None = None,
                    U_prev: torch.Tensor | None = None) -> dict:
    """One K3/K4 launch over an element block (asynchronous on ``stream``).
    ``outputs`` ⊆ {xi, R_elem, K_elem, R_global, sigma, iters, flags}; ``R_global``
    (atomic scatter-add) is zero-initialised here unless passed in through ``out``.
    ``U_prev``: the previous step's displacement vector, required by
    ``small_rate_elastic_plastic`` blocks (their residual sees eps(U) - eps(U_prev))."""
    n_e, n_b, n_ip = arrays.n_elems, arrays.n_basis, arrays.n_ip
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    n_xi = 7
    if U_global.dtype != torch.float64 or U_global.numel() != arrays.n_dofs or not U_global.is_contiguous():
        raise ValueError(f"U_global: expected contiguous float64 ({arrays.n_dofs},)")
    if xi_prev.dtype != torch.float64 or tuple(xi_prev.shape) != (n_e, n_ip, n_xi) or not xi_prev.is_contiguous():
        raise ValueError(f"xi_prev: expected contiguous float64 ({n_e}, {n_ip}, {n_xi}), got {tuple(xi_prev.shape)}")
    for t in (arrays.elem_eq, arrays.grad_N, arrays.det, arrays.quad_w):
        if not t.is_contiguous():
            raise ValueError("FE block arrays must be contiguous")
    shapes = {"xi": ((n_e, n_ip, n_xi), torch.float64), "R_elem": ((n_e, n_b * 3), torch.float64),
              "K_elem": ((n_e, n_b * 3, n_b * 3), torch.float64), "R_global": ((arrays.n_dofs,), torch.float64),
              "sigma": ((n_e, n_ip, 6), torch.float64), "iters": ((n_e, n_ip), torch.int32),
              "flags": ((n_e, n_ip), torch.int32)}
    out = dict(out) if out is not None else {}
    for name in set(outputs) | {"xi"}:
        if name not in shapes:
            raise ValueError(f"unknown output {name!r}")
        if out.get(name) is None:
            shape, dt = shapes[name]
            alloc = torch.zeros if name == "R_global" else torch.empty
            out[name] = alloc(shape, dtype=dt, device=dev)
    if U_prev is not None and (U_prev.dtype != torch.float64 or U_prev.numel() != arrays.n_dofs
                               or not U_prev.is_contiguous() or U_prev.device != dev):
        raise ValueError(f"U_prev: expected contiguous float64 ({arrays.n_dofs},) on {dev}")
    b = _fe_struct(arrays, U_global, xi_prev, out, U_prev)
    nw = newton.to_struct()
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = L.lib().cmadx_fe_block_assemble(C.byref(material), C.byref(nw), C.byref(b),
                                             C.c_void_p(s.cuda_stream))
    L.check(rc, "cmadx_fe_block_assemble")
    return out


def assemble_element_block(material: L.Material, newton: NewtonSettings, arrays: FEBlockArrays,
                           U_global: torch.Tensor, xi_prev_per_block: torch.Tensor,
                           r_plan: SegmentPlan | None = None, out: dict | None = None,
                           stream: torch.cuda.Stream | None = None, U_prev: torch.Tensor | None = None):
    """``(R_block, vals, xi_solved_per_block)`` of a COUPLED block, as the
    reference's ``assemble_element_block`` returns them: ``R_block (n_dofs,)``,
    ``vals`` = flattened ``(elem, row dof, col dof)`` COO data,
    ``xi_solved (n_e, n_ip, n_xi)``.  With ``r_plan`` (a :class:`SegmentPlan` over
    ``elem_eq.ravel()``) R is summed deterministically; otherwise by atomics."""
    outs = ("xi", "R_elem", "K_elem") if r_plan is not None else ("xi", "K_elem", "R_global")
    o = fe_block_launch(material, newton, arrays, U_global, xi_prev_per_block, outs, out, stream, U_prev=U_prev)
    R = r_plan.sum(o["R_elem"].reshape(-1), stream=stream) if r_plan is not None else o["R_global"]
    return R, o["K_elem"].reshape(-1), o["xi"]


def assemble_element_block_residual(material, newton, arrays, U_global, xi_prev_per_block,
                                    r_plan: SegmentPlan | None = None, out: dict | None = None,
                                    stream=None) -> torch.Tensor:
    """Residual-only block assembly (K4): line-search probes and reaction QoIs
    (cmad/fem/assembly.py:735-813, cmad/qois/fe_load_match.py:179-196)."""
    outs = ("xi", "R_elem") if r_plan is not None else ("xi", "R_global")
    o = fe_block_launch(material, newton, arrays, U_global, xi_prev_per_block, outs, out, stream)
    return r_plan.sum(o["R_elem"].reshape(-1), stream=stream) if r_plan is not None else o["R_global"]


def assemble_element_block_mixed(material: L.Material, newton: NewtonSettings, arrays: FEBlockArrays,
                                 U_global: torch.Tensor, xi_prev_per_block: torch.Tensor,
                                 stab_mult: float = 1.0, r_plan: SegmentPlan | None = None,
                                 want_K: bool = True, stream: torch.cuda.Stream | None = None,
                                 U_prev: torch.Tensor | None = None):
    """Mixed u-p (``SmallDispEquilibrium(mixed=True)``, small_disp_equilibrium.py:87-111)
    counterpart of :func:`assemble_element_block`: returns ``(R_block, vals, xi_solved)``
    with ``R_block (n_dofs,)`` over the block-major (u, p) dofs and ``vals`` the
    concatenated COO streams in the reference's (r, s) emit order - (u,u), (u,p), (p,u),
    (p,p), each flattened ``(elem, row dof, col dof)`` (cmad/fem/assembly.py:722-732).
    ``r_plan``: a :class:`SegmentPlan` over ``cat(elem_eq.ravel(), elem_eq_p.ravel())``
    (deterministic R); otherwise atomics.  ``want_K=False``: residual only (vals is None).
    ``U_prev``: the previous step's (u, p) vector, required by ``small_rate_elastic_plastic`` blocks -
    their pressure rows depend on the local state (``hydro_cauchy = tr(cauchy(xi)) / 3``) and are
    formed by the element kernel itself."""
    if not arrays.mixed:
        raise ValueError("block arrays were built without mixed=True")
    n_e, n_b, n_ip = arrays.n_elems, arrays.n_basis, arrays.n_ip
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    if U_global.dtype != torch.float64 or U_global.numel() != arrays.n_dofs or not U_global.is_contiguous():
        raise ValueError(f"U_global: expected contiguous float64 ({arrays.n_dofs},)")
    if xi_prev_per_block.dtype != torch.float64 or tuple(xi_prev_per_block.shape) != (n_e, n_ip, 7) \
            or not xi_prev_per_block.is_contiguous():
        raise ValueError(f"xi_prev: expected contiguous float64 ({n_e}, {n_ip}, 7)")
    nu, npd = 3 * n_b, n_b
    sizes = [n_e * nu * nu, n_e * nu * npd, n_e * npd * nu, n_e * npd * npd]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    vals = torch.empty(int(offs[-1]), dtype=torch.float64, device=dev) if want_K else None
    R_elem = torch.empty(n_e * (nu + npd), dtype=torch.float64, device=dev)     # [R_u | R_p]
    out = {"xi": torch.empty((n_e, n_ip, 7), dtype=torch.float64, device=dev),
           "R_elem": R_elem[:n_e * nu].view(n_e, nu),
           "K_elem": vals[:sizes[0]].view(n_e, nu, nu) if want_K else None}
    R = None
    if r_plan is None:
        R = torch.zeros(arrays.n_dofs, dtype=torch.float64, device=dev)
        out["R_global"] = R
    if U_prev is not None and (U_prev.dtype != torch.float64 or U_prev.numel() != arrays.n_dofs
                               or not U_prev.is_contiguous() or U_prev.device != dev):
        raise ValueError(f"U_prev: expected contiguous float64 ({arrays.n_dofs},) on {dev}")
    b = _fe_struct(arrays, U_global, xi_prev_per_block, out, U_prev)
    mx = L.FeMixed()
    mx.elem_eq_p, mx.N, mx.h = arrays.elem_eq_p.data_ptr(), arrays.N.data_ptr(), arrays.h.data_ptr()
    mx.stab_mult = float(stab_mult)
    mx.R_p_elem = R_elem[n_e * nu:].data_ptr()
    if want_K:
        mx.K_up, mx.K_pu, mx.K_pp = (vals[int(offs[i]):].data_ptr() for i in (1, 2, 3))
    mx.R_global = R.data_ptr() if R is not None else None
    nw = newton.to_struct()
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = L.lib().cmadx_fe_block_assemble_mixed(C.byref(material), C.byref(nw), C.byref(b), C.byref(mx),
                                                   C.c_void_p(s.cuda_stream))
    L.check(rc, "cmadx_fe_block_assemble_mixed")
    if r_plan is not None:
        R = r_plan.sum(R_elem, stream=stream)
    return R, vals, out["xi"]


def mixed_r_plan(arrays: FEBlockArrays, device=None) -> SegmentPlan:
    """Deterministic scatter plan of the mixed block's residual: items = ``[R_u | R_p]``."""
    seg = np.concatenate([arrays.elem_eq.cpu().numpy().reshape(-1), arrays.elem_eq_p.cpu().numpy().reshape(-1)])
    return SegmentPlan(seg, arrays.n_dofs, device=device if device is not None else arrays.grad_N.device)


def evaluate_cauchy_at_ips(material: L.Material, arrays: FEBlockArrays, U_global: torch.Tensor,
                           xi: torch.Tensor, out: torch.Tensor | None = None,
                           stream: torch.cuda.Stream | None = None) -> torch.Tensor:
    """``(n_elems, n_ip, 6)`` cauchy stress at every integration point of a COUPLED block from
    the stored converged state ``xi`` (cmad/fem/postprocess.py:35-185, ``model.cauchy`` branch):
    no Newton, one HBM-bound launch."""
    n_e, n_ip = arrays.n_elems, arrays.n_ip
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    if U_global.dtype != torch.float64 or U_global.numel() != arrays.n_dofs or not U_global.is_contiguous():
        raise ValueError(f"U_global: expected contiguous float64 ({arrays.n_dofs},)")
    if xi.dtype != torch.float64 or tuple(xi.shape) != (n_e, n_ip, 7) or not xi.is_contiguous():
        raise ValueError(f"xi: expected contiguous float64 ({n_e}, {n_ip}, 7)")
    if out is None:
        out = torch.empty((n_e, n_ip, 6), dtype=torch.float64, device=dev)
    b = _fe_struct(arrays, U_global, xi, {"xi": out})
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = L.lib().cmadx_fe_cauchy_at_ips(C.byref(material), C.byref(b), C.c_void_p(xi.data_ptr()),
                                            C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream))
    L.check(rc, "cmadx_fe_cauchy_at_ips")
    return out


class ReactionPlan:
    """``FELoadMatch._reaction_at`` (cmad/qois/fe_load_match.py:179-196): the assembled residual
    summed over each requested component's sideset dofs, deterministically (K5 over the
    gathered entries).  ``eq_per_component``: list of int arrays of global equations."""

    def __init__(self, eq_per_component, device):
        self.device = torch.device(device)
        eqs = [np.asarray(e, dtype=np.int64).reshape(-1) for e in eq_per_component]
        self._gather = torch.from_numpy(np.concatenate(eqs) if eqs else np.zeros(0, np.int64)).to(self.device)
        seg = np.concatenate([np.full(len(e), c, dtype=np.int64) for c, e in enumerate(eqs)]) if eqs \
            else np.zeros(0, np.int64)
        self._plan = SegmentPlan(seg, len(eqs), device=self.device)

    def __call__(self, R: torch.Tensor) -> torch.Tensor:
        return self._plan.sum(R.index_select(0, self._gather).contiguous())


class EmbeddedBCPlan:
    """Device-side ``_embedded_bc_enforce`` + ``_embedded_residual``
    (cmad/fem/sparse_solve.py:1058-1174) over the deduplicated COO pattern ``(rows, cols)``
    for a fixed prescribed set: ``apply(K_data, R, U, presc_vals) -> (r, K_emb_data)``."""

    def __init__(self, rows, cols, n_dofs: int, presc_idx, device=None):
        rows = np.ascontiguousarray(rows, dtype=np.int64); cols = np.ascontiguousarray(cols, dtype=np.int64)
        idx = np.ascontiguousarray(presc_idx, dtype=np.int64)
        self.n, self.nnz, self.n_presc = int(n_dofs), int(rows.size), int(idx.size)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        p64 = C.POINTER(C.c_int64)
        with torch.cuda.device(self.device):
            rc = L.lib().cmadx_embedded_plan_create(rows.ctypes.data_as(p64), cols.ctypes.data_as(p64), self.nnz,
                                                    self.n, idx.ctypes.data_as(p64), self.n_presc, C.byref(self._h))
        L.check(rc, "cmadx_embedded_plan_create")

    def apply(self, K_data: torch.Tensor, R: torch.Tensor, U: torch.Tensor, presc_vals: torch.Tensor,
              stream: torch.cuda.Stream | None = None):
        for name, t, n in (("K_data", K_data, self.nnz), ("R", R, self.n), ("U", U, self.n),
                           ("presc_vals", presc_vals, self.n_presc)):
            if t.dtype != torch.float64 or t.numel() != n or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"{name}: expected contiguous float64 with {n} entries on {self.device}")
        r = torch.empty(self.n, dtype=torch.float64, device=self.device)
        K_emb = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            rc = L.lib().cmadx_embedded_apply(self._h, K_data.data_ptr(), R.data_ptr(), U.data_ptr(),
                                              presc_vals.data_ptr(), r.data_ptr(), K_emb.data_ptr(), s.cuda_stream)
        L.check(rc, "cmadx_embedded_apply")
        return r, K_emb

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().cmadx_embedded_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fe_block_jvp(material: L.Material, arrays: FEBlockArrays, U_global: torch.Tensor,
                 xi_prev: torch.Tensor, xi_state: torch.Tensor, active_pid, dp,
                 dxi_prev: torch.Tensor | None = None, outputs=("xi", "R_elem"), out: dict | None = None,
                 stream: torch.cuda.Stream | None = None, dU: torch.Tensor | None = None,
                 stab_mult: float | None = None) -> dict:
    """K6: tangent of the converged block w.r.t. (params, xi_prev) at fixed ``U``:
    returns ``out["xi"]`` = ``dxi (n_e, n_ip, 7)`` and ``out["R_elem"]`` = ``dR_e``
    (or ``R_global``) for the direction ``dp`` (native values of the active
    parameters, in ``active_pid`` order) and ``dxi_prev``.  What ``jax.jvp`` pushes
    through the assembled residual inside the FE Newton's IFT rule
    (cmad/fem/nonlinear_solver.py:490-537).  ``dU (n_dofs,)`` optionally adds a displacement
    direction (then ``dR`` includes ``K dU`` and ``dxi`` is the total state sensitivity).
    ``stab_mult`` (mixed u-p block arrays only) selects the mixed formulation: ``U`` / ``dU``
    cover the block-major (u, p) dofs, ``R_elem`` is ``dR_u`` and the extra output ``R_p_elem
    (n_e, n_b)`` is ``dR_p``; ``R_global`` accumulates both."""
    mixed = stab_mult is not None
    if mixed and not arrays.mixed:
        raise ValueError("block arrays were built without mixed=True")
    n_e, n_b, n_ip = arrays.n_elems, arrays.n_basis, arrays.n_ip
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    if dU is not None and (dU.dtype != torch.float64 or dU.numel() != arrays.n_dofs or not dU.is_contiguous()):
        raise ValueError(f"dU: expected contiguous float64 ({arrays.n_dofs},)")
    for name, t in (("xi_prev", xi_prev), ("xi_state", xi_state), ("dxi_prev", dxi_prev)):
        if t is None:
            continue
        if t.dtype != torch.float64 or tuple(t.shape) != (n_e, n_ip, 7) or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous float64 ({n_e}, {n_ip}, 7)")
    if U_global.dtype != torch.float64 or U_global.numel() != arrays.n_dofs or not U_global.is_contiguous():
        raise ValueError(f"U_global: expected contiguous float64 ({arrays.n_dofs},)")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    dpv = np.ascontiguousarray(dp, dtype=np.float64)
    if pid.shape != dpv.shape or pid.ndim != 1:
        raise ValueError("dp must have one entry per active parameter")
    shapes = {"xi": ((n_e, n_ip, 7), torch.float64), "R_elem": ((n_e, n_b * 3), torch.float64),
              "R_global": ((arrays.n_dofs,), torch.float64)}
    if mixed:
        shapes["R_p_elem"] = ((n_e, n_b), torch.float64)
        outputs = tuple(outputs) + (("R_p_elem",) if "R_elem" in outputs else ())
    out = dict(out) if out is not None else {}
    for name in set(outputs) | {"xi"}:
        if name not in shapes:
            raise ValueError(f"unknown JVP output {name!r}")
        if out.get(name) is None:
            shape, dt = shapes[name]
            out[name] = (torch.zeros if name == "R_global" else torch.empty)(shape, dtype=dt, device=dev)
    b = _fe_struct(arrays, U_global, xi_prev, out)
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    tail = (C.c_void_p(xi_state.data_ptr()),
            C.c_void_p(dxi_prev.data_ptr()) if dxi_prev is not None else None,
            C.c_void_p(dU.data_ptr()) if dU is not None else None, C.c_void_p(s.cuda_stream))
    with torch.cuda.device(dev):
        if mixed:
            mx = _fe_mixed_struct(arrays, stab_mult)
            if out.get("R_p_elem") is not None:
                mx.R_p_elem = out["R_p_elem"].data_ptr()
            if out.get("R_global") is not None:
                mx.R_global = out["R_global"].data_ptr()
            rc = L.lib().cmadx_fe_block_jvp_mixed(
                C.byref(material), pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                dpv.ctypes.data_as(C.POINTER(C.c_double)), C.byref(b), C.byref(mx), *tail)
        else:
            rc = L.lib().cmadx_fe_block_jvp(
                C.byref(material), pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                dpv.ctypes.data_as(C.POINTER(C.c_double)), C.byref(b), *tail)
    L.check(rc, "cmadx_fe_block_jvp")
    return out


def _fe_mixed_struct(arrays: FEBlockArrays, stab_mult: float) -> "L.FeMixed":
    mx = L.FeMixed()
    mx.elem_eq_p, mx.N, mx.h = arrays.elem_eq_p.data_ptr(), arrays.N.data_ptr(), arrays.h.data_ptr()
    mx.stab_mult = float(stab_mult)
    return mx


def fe_block_vjp(material: L.Material, arrays: FEBlockArrays, U_global: torch.Tensor,
                 xi_prev: torch.Tensor, xi_state: torch.Tensor, active_pid, Rbar: torch.Tensor,
                 xibar: torch.Tensor | None = None, group=None,
                 stream: torch.cuda.Stream | None = None,
                 stab_mult: float | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """K6 reverse mode: ``(pbar (n_active,), xibar_prev (n_e, n_ip, 7))`` for the cotangents
    ``Rbar (n_dofs,)`` of the assembled residual and ``xibar`` of the converged local
    state - one step of a discrete FE adjoint, the transpose of :func:`fe_block_jvp`.
    With ``group`` given (``comm.WORLD`` or a process group: the elements are this rank's
    partition) ``pbar`` is all-reduced: the gradient exchange of the calibration loop;
    ``group=None`` never communicates.  ``stab_mult`` (mixed block arrays) selects the mixed
    u-p formulation: ``Rbar`` covers both residual blocks."""
    mixed = stab_mult is not None
    if mixed and not arrays.mixed:
        raise ValueError("block arrays were built without mixed=True")
    n_e, n_ip = arrays.n_elems, arrays.n_ip
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    for name, t in (("xi_prev", xi_prev), ("xi_state", xi_state), ("xibar", xibar)):
        if t is not None and (t.dtype != torch.float64 or tuple(t.shape) != (n_e, n_ip, 7) or not t.is_contiguous()):
            raise ValueError(f"{name}: expected contiguous float64 ({n_e}, {n_ip}, 7)")
    for name, t in (("U_global", U_global), ("Rbar", Rbar)):
        if t.dtype != torch.float64 or t.numel() != arrays.n_dofs or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous float64 ({arrays.n_dofs},)")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    na = len(pid)
    out = {"xi": torch.empty((n_e, n_ip, 7), dtype=torch.float64, device=dev)}
    b = _fe_struct(arrays, U_global, xi_prev, out)
    pbar = torch.zeros((max(na, 1),), dtype=torch.float64, device=dev)
    wsb = int(L.lib().cmadx_fe_vjp_workspace_bytes(n_e, n_ip, na))
    ws = torch.empty((max(wsb // 8, 1),), dtype=torch.float64, device=dev)
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    tail = (C.c_void_p(xi_state.data_ptr()), C.c_void_p(Rbar.data_ptr()),
            C.c_void_p(xibar.data_ptr()) if xibar is not None else None,
            C.c_void_p(pbar.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(s.cuda_stream))
    with torch.cuda.device(dev):
        if mixed:
            mx = _fe_mixed_struct(arrays, stab_mult)
            rc = L.lib().cmadx_fe_block_vjp_mixed(
                C.byref(material), pid.ctypes.data_as(C.POINTER(C.c_int32)), na, C.byref(b), C.byref(mx), *tail)
        else:
            rc = L.lib().cmadx_fe_block_vjp(
                C.byref(material), pid.ctypes.data_as(C.POINTER(C.c_int32)), na, C.byref(b), *tail)
    L.check(rc, "cmadx_fe_block_vjp")
    pbar = all_reduce_sum(pbar[:na], group)
    return pbar, out["xi"]


def fe_block_vjp_disp(material: L.Material, arrays: FEBlockArrays, U_global: torch.Tensor,
                      xi_prev: torch.Tensor, xi_state: torch.Tensor, xibar: torch.Tensor | None = None,
                      Rbar: torch.Tensor | None = None, stream: torch.cuda.Stream | None = None,
                      stab_mult: float | None = None) -> torch.Tensor:
    """Displacement cotangent of the converged block, per integration point:
    ``Ubar_ip (n_e, n_ip, 3 n_b)`` with ``sum_ip`` scattered over ``elem_eq`` =
    ``(d xi/dU)^T xibar + (d R_u/dU |total)^T Rbar`` - the transpose of the displacement
    direction of :func:`fe_block_jvp`, the piece a discrete FE adjoint through the load steps
    needs besides the assembled tangent (:func:`cmad_b200.fe_driver.fe_adjoint_gradient`)."""
    n_e, n_ip, n_b = arrays.n_elems, arrays.n_ip, arrays.n_basis
    dev = arrays.grad_N.device
    if dev.type != "cuda":
        raise ValueError("FE block arrays must live on a CUDA device (there is no CPU fallback)")
    for name, t in (("xi_prev", xi_prev), ("xi_state", xi_state), ("xibar", xibar)):
        if t is not None and (t.dtype != torch.float64 or tuple(t.shape) != (n_e, n_ip, 7) or not t.is_contiguous()):
            raise ValueError(f"{name}: expected contiguous float64 ({n_e}, {n_ip}, 7)")
    for name, t in (("U_global", U_global), ("Rbar", Rbar)):
        if t is not None and (t.dtype != torch.float64 or t.numel() != arrays.n_dofs or not t.is_contiguous()):
            raise ValueError(f"{name}: expected contiguous float64 ({arrays.n_dofs},)")
    out = {"xi": torch.empty((n_e, n_ip, 7), dtype=torch.float64, device=dev)}
    b = _fe_struct(arrays, U_global, xi_prev, out)
    Ubar_ip = torch.empty((n_e, n_ip, 3 * n_b), dtype=torch.float64, device=dev)
    mx = _fe_mixed_struct(arrays, stab_mult) if stab_mult is not None else None
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = L.lib().cmadx_fe_block_vjp_disp(
            C.byref(material), C.byref(b), C.byref(mx) if mx is not None else None,
            C.c_void_p(xi_state.data_ptr()), C.c_void_p(Rbar.data_ptr()) if Rbar is not None else None,
            C.c_void_p(xibar.data_ptr()) if xibar is not None else None,
            C.c_void_p(Ubar_ip.data_ptr()), C.c_void_p(s.cuda_stream))
    L.check(rc, "cmadx_fe_block_vjp_disp")
    return Ubar_ip


def disp_cotangent_plan(arrays: FEBlockArrays, device=None) -> SegmentPlan:
    """Deterministic scatter plan of :func:`fe_block_vjp_disp`'s per-point rows into the nodal
    vector: items = the element's equation row repeated for each of its points."""
    eq = arrays.elem_eq.cpu().numpy()
    seg = np.repeat(eq[:, None, :], arrays.n_ip, axis=1).reshape(-1)
    return SegmentPlan(seg, arrays.n_dofs, device=device if device is not None else arrays.grad_N.device)


def partition_block(arrays: FEBlockArrays, rank: int, world: int) -> tuple[FEBlockArrays, tuple[int, int]]:
    """This rank's contiguous element range of a block (the path shards by element:
    every element's local state, K_e and R_e are computed by exactly one rank)."""
    from .objectives import shard_range
    lo, hi = shard_range(arrays.n_elems, rank, world)
    return arrays.slice(lo, hi), (lo, hi)


def reduce_residual(R_local: torch.Tensor, group=None) -> torch.Tensor:
    """The one exchange step of the FE path: sum the per-rank scatter-added residuals
    (shared-node entries get contributions from several ranks).  NCCL over NVLink on
    GPUs, gloo in the CPU tests; opt-in (``group=comm.WORLD`` or a process group), a no-op
    with ``group=None``."""
    return all_reduce_sum(R_local, group)


class InterfaceExchange:
    """Halo exchange of the assembled residual for an element partition: only the dofs touched
    by MORE than one rank need summing, so the per-step collective is one all-reduce of the
    packed interface entries (O(interface) bytes) instead of the whole vector (O(n_dofs)).
    After :meth:`reduce`, ``R`` is complete on every dof this rank's elements touch (its own
    interior + the interfaces it shares); entries of other ranks' interiors stay untouched.
    Set-up costs one all-reduce of an int32 touch-count vector, once per mesh partition.
    NCCL over NVLink on GPUs; the same code runs on gloo in the CPU tests."""

    def __init__(self, elem_eq: torch.Tensor, n_dofs: int, group=None, extra_eq: torch.Tensor | None = None):
        import torch.distributed as dist
        self._active, self._group = comm_resolve(group)
        dev = elem_eq.device
        touched = torch.zeros(n_dofs, dtype=torch.int32, device=dev)
        eqs = [elem_eq.reshape(-1).long()] + ([extra_eq.reshape(-1).long()] if extra_eq is not None else [])
        for e in eqs:
            touched[e] = 1
        self.mine = touched.bool()
        count = touched.clone()
        if self._active:
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=self._group)
        self.index = torch.nonzero(count > 1).reshape(-1)          # same on every rank, sorted
        self.n_interface = int(self.index.numel())
        self._buf = torch.empty(self.n_interface, dtype=torch.float64, device=dev)

    def reduce(self, R: torch.Tensor) -> torch.Tensor:
        if not self._active or self.n_interface == 0:
            return R
        import torch.distributed as dist
        if R.is_cuda:       # pack / unpack with the library's own index kernels (no framework kernels on the step path)
            s = C.c_void_p(torch.cuda.current_stream(R.device).cuda_stream)
            with torch.cuda.device(R.device):
                L.check(L.lib().cmadx_index_gather(self.index.data_ptr(), self.n_interface, R.data_ptr(),
                                                   self._buf.data_ptr(), s), "cmadx_index_gather")
                dist.all_reduce(self._buf, op=dist.ReduceOp.SUM, group=self._group)
                L.check(L.lib().cmadx_index_scatter(self.index.data_ptr(), self.n_interface, self._buf.data_ptr(),
                                                    R.data_ptr(), s), "cmadx_index_scatter")
            return R
        torch.index_select(R, 0, self.index, out=self._buf)           # CPU tensors (gloo tests of the host logic)
        dist.all_reduce(self._buf, op=dist.ReduceOp.SUM, group=self._group)
        R.index_copy_(0, self.index, self._buf)
        return R


def assemble_global(blocks: Mapping[str, tuple], U_global: torch.Tensor,
                    xi_prev_by_block: Mapping[str, torch.Tensor], coo_plan: SegmentPlan | None = None,
                    r_plans: Mapping[str, SegmentPlan] | None = None, group=None):
    """Walk all element blocks (``blocks[name] = (material, newton, arrays)``) and
    return ``(K_data, R, xi_solved_by_block)``: ``K_data`` is the deduplicated COO
    data (``coo_plan`` over the concatenated ``coo_dedup_scatter``) or, without a
    plan, the with-duplicates stream; ``R`` the global residual.  Under
    ``torch.distributed`` the element blocks passed in are this rank's partition
    (:func:`partition_block`) and ``R`` is all-reduced - the one exchange step of the
    path; ``K_data`` and ``xi`` stay rank-local (element-owned)."""
    R = torch.zeros(U_global.numel(), dtype=torch.float64, device=U_global.device)
    vals_all, xi_out = [], {}
    for name, (material, newton, arrays) in blocks.items():
        plan = r_plans.get(name) if r_plans else None
        Rb, vals, xi = assemble_element_block(material, newton, arrays, U_global,
                                              xi_prev_by_block[name], r_plan=plan)
        R += Rb
        vals_all.append(vals)
        xi_out[name] = xi
    R = reduce_residual(R, group)
    vals = vals_all[0] if len(vals_all) == 1 else torch.cat(vals_all)
    K = coo_plan.sum(vals) if coo_plan is not None else vals
    return K, R, xi_out
