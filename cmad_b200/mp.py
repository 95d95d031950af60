"""Batched material-point update: Python front end of K1 (``cmadx_mp_update``).

Arrays are torch CUDA tensors in the component-major ("SoA") layout of the
C-ABI: ``xi_prev (n_xi, N)``, ``strain (6|9, N)``; torch is used only to own
device memory and streams.  The host-buffer variant takes NumPy / pinned
torch CPU tensors and runs the library's chunked H2D / kernel / D2H pipeline.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .material import NewtonSettings

OUTPUT_COMPONENTS = {  # name -> components as a function of (n_xi, n_active, n_strain_dirs)
    "xi": lambda n, a, s=6: n, "sigma": lambda n, a, s=6: 6, "dsig_deps": lambda n, a, s=6: 6 * s,
    "dxi_deps": lambda n, a, s=6: n * s, "dC_dp": lambda n, a, s=6: n * a, "dC_dxi": lambda n, a, s=6: n * n,
    "dC_dxi_prev": lambda n, a, s=6: n * n, "C": lambda n, a, s=6: n,
}
# deformation types (cmad/models/deformation_types.py): n_xi, prescribed symmetric strain
# components (= derivative directions), accepted strain rows
DEF_TYPES = {L.DEF_FULL_3D: (7, 6, (6, 9)), L.DEF_PLANE_STRESS: (8, 3, (3, 4)), L.DEF_UNIAXIAL_STRESS: (9, 1, (1,))}
POINT_SCALARS = {"iters": torch.int32, "flags": torch.int32, "cnorm": torch.float64}
DEFAULT_OUTPUTS = ("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags")


def n_xi_of(material: L.Material, def_type: int = L.DEF_FULL_3D) -> int:
    if def_type != L.DEF_FULL_3D:
        # the rate form under uniaxial stress carries three off-axis delta strains more
        # (small_rate_elastic_plastic.py:189-199)
        extra = 3 if (material.model == L.MODEL_SMALL_RATE_ELASTIC_PLASTIC and def_type == L.DEF_UNIAXIAL_STRESS) else 0
        return DEF_TYPES[def_type][0] + extra
    return 6 if material.model == L.MODEL_ELASTIC else 7


def init_xi(material: L.Material, n: int, device, def_type: int = L.DEF_FULL_3D) -> torch.Tensor:
    """Initial local state (small_elastic_plastic.py:139-180): zeros, stretches = 1."""
    nxi = n_xi_of(material, def_type)
    x = torch.zeros((nxi, n), dtype=torch.float64, device=device)
    if def_type != L.DEF_FULL_3D:
        x[7:DEF_TYPES[def_type][0]] = 1.0          # the stretches; the rate form's delta strains stay 0
    return x


def _check_in(name, t, rows, n, device_type):
    if t.dtype != torch.float64 or t.dim() != 2 or t.shape[0] != rows or t.shape[1] != n:
        raise ValueError(f"{name}: expected float64 ({rows}, {n}), got {t.dtype} {tuple(t.shape)}")
    if t.device.type != device_type:
        raise ValueError(f"{name}: expected a {device_type} tensor")
    if t.shape[1] > 1 and t.stride(1) != 1:          # (a size-1 axis may carry any stride)
        raise ValueError(f"{name}: innermost (point) dimension must be contiguous")


def allocate_outputs(material: L.Material, n: int, n_active: int, outputs, device,
                     pin: bool = False, def_type: int = L.DEF_FULL_3D) -> dict:
    nxi = n_xi_of(material, def_type)
    ns = DEF_TYPES[def_type][1]
    out = {}
    kw = dict(device=device)
    if pin:
        kw = dict(device="cpu", pin_memory=True)
    for name in outputs:
        if name in OUTPUT_COMPONENTS:
            rows = OUTPUT_COMPONENTS[name](nxi, n_active, ns)
            if name == "dC_dp" and n_active == 0:
                continue
            out[name] = torch.empty((rows, n), dtype=torch.float64, **kw)
        elif name in POINT_SCALARS:
            out[name] = torch.empty((n,), dtype=POINT_SCALARS[name], **kw)
        else:
            raise ValueError(f"unknown output {name!r}")
    return out


def _buffers(material, xi_prev, strain, xi_init, out: dict, def_type: int = L.DEF_FULL_3D) -> L.MpBuffers:
    n = xi_prev.shape[1]
    ld = xi_prev.stride(0) if xi_prev.shape[0] > 1 else max(n, 1)
    tensors = [xi_prev, strain] + ([xi_init] if xi_init is not None else []) + \
        [t for k, t in out.items() if k in OUTPUT_COMPONENTS]
    for t in tensors:
        s0 = t.stride(0) if t.shape[0] > 1 else ld
        if s0 != ld:
            raise ValueError("all component-major arrays must share one leading dimension")
    b = L.MpBuffers()
    b.n, b.ld, b.strain_comps, b.def_type = n, ld, strain.shape[0], int(def_type)
    b.xi_prev, b.strain = xi_prev.data_ptr(), strain.data_ptr()
    b.xi_init = xi_init.data_ptr() if xi_init is not None else None
    for name in list(OUTPUT_COMPONENTS) + list(POINT_SCALARS):
        setattr(b, name, out[name].data_ptr() if name in out else None)
    return b


def mp_update(material: L.Material, newton: NewtonSettings, active_pid, xi_prev: torch.Tensor,
              strain: torch.Tensor, outputs=DEFAULT_OUTPUTS, out: dict | None = None,
              xi_init: torch.Tensor | None = None, stream: torch.cuda.Stream | None = None,
              def_type: int = L.DEF_FULL_3D) -> dict:
    """One batched constitutive update on the GPU (asynchronous on ``stream``).

    Returns the dict of requested output tensors (allocated unless ``out`` is
    given).  Raises if the CUDA library is unavailable - there is no fallback.
    ``def_type``: FULL_3D (default), PLANE_STRESS (``xi`` has 8 rows, ``strain`` 3 =
    (e_xx, e_xy, e_yy) or 4 = 2x2 grad_u) or UNIAXIAL_STRESS (9 rows, 1 strain row).
    """
    lib = L.lib()
    if def_type not in DEF_TYPES:
        raise ValueError(f"unknown def_type {def_type}")
    nxi = n_xi_of(material, def_type)
    n = xi_prev.shape[1]
    _check_in("xi_prev", xi_prev, nxi, n, "cuda")
    if strain.shape[0] not in DEF_TYPES[def_type][2]:
        raise ValueError(f"strain must have one of {DEF_TYPES[def_type][2]} components for this def_type")
    _check_in("strain", strain, strain.shape[0], n, "cuda")
    if xi_init is not None:
        _check_in("xi_init", xi_init, nxi, n, "cuda")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    if out is None:
        out = allocate_outputs(material, n, len(pid), outputs, xi_prev.device, def_type=def_type)
    b = _buffers(material, xi_prev, strain, xi_init, out, def_type)
    nw = newton.to_struct()
    s = stream if stream is not None else torch.cuda.current_stream(xi_prev.device)
    with torch.cuda.device(xi_prev.device):
        rc = lib.cmadx_mp_update(C.byref(material), C.byref(nw),
                                 pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                                 C.byref(b), C.c_void_p(s.cuda_stream))
    L.check(rc, "cmadx_mp_update")
    return out


def mp_update_host(material: L.Material, newton: NewtonSettings, active_pid, xi_prev, strain,
                   outputs=DEFAULT_OUTPUTS, out: dict | None = None, xi_init=None,
                   device: int = 0, chunk_points: int = 0, def_type: int = L.DEF_FULL_3D) -> dict:
    """Same update on HOST buffers (NumPy arrays or CPU torch tensors, ideally
    pinned): chunked H2D -> kernel -> D2H pipeline inside the library. Blocking.
    Every deformation type of :func:`mp_update`."""
    lib = L.lib()
    if def_type not in DEF_TYPES:
        raise ValueError(f"unknown def_type {def_type}")
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    xi_prev, strain = as_t(xi_prev), as_t(strain)
    xi_init = as_t(xi_init) if xi_init is not None else None
    nxi = n_xi_of(material, def_type)
    n = xi_prev.shape[1]
    _check_in("xi_prev", xi_prev, nxi, n, "cpu")
    if strain.shape[0] not in DEF_TYPES[def_type][2]:
        raise ValueError(f"strain must have one of {DEF_TYPES[def_type][2]} components for this def_type")
    _check_in("strain", strain, strain.shape[0], n, "cpu")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    if out is None:
        out = allocate_outputs(material, n, len(pid), outputs, "cpu", pin=False, def_type=def_type)
    b = _buffers(material, xi_prev, strain, xi_init, out, def_type)
    nw = newton.to_struct()
    rc = lib.cmadx_mp_update_host(C.byref(material), C.byref(nw),
                                  pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                                  C.byref(b), int(device), int(chunk_points))
    L.check(rc, "cmadx_mp_update_host")
    return out


def mp_objective_host(material: L.Material, newton: NewtonSettings, active_pid, strain_hist, data_hist, weight,
                      strategy: str = "adjoint", device: int = 0, chunk_points: int = 0,
                      want_J_point: bool = False):
    """Calibration objective ``(J, dJ/dp native)`` of a batch of experiments held in HOST memory
    (``cmadx_mp_objective_host``): ``strain_hist (N+1, comps, n)``, ``data_hist (N+1, 9, n)`` NumPy /
    CPU tensors (pinned for speed), ``weight`` 3x3.  Returns ``result (1 + n_active,)`` and, with
    ``want_J_point``, the per-point objective.  Histories go to the device in chunks of points;
    what comes back is 8 (1 + n_active) bytes per chunk."""
    lib = L.lib()
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    sh, dh = as_t(strain_hist), as_t(data_hist)
    if sh.dtype != torch.float64 or dh.dtype != torch.float64 or sh.dim() != 3 or dh.dim() != 3 or not sh.is_contiguous() \
            or not dh.is_contiguous() or dh.shape[1] != 9 or dh.shape[0] != sh.shape[0] or dh.shape[2] != sh.shape[2]:
        raise ValueError("strain_hist (N+1, comps, n) and data_hist (N+1, 9, n): contiguous float64")
    if sh.device.type != "cpu" or dh.device.type != "cpu":
        raise ValueError("mp_objective_host takes host arrays (device histories: cmad_b200.objectives)")
    if strategy not in ("adjoint", "direct"):
        raise ValueError(f"unknown strategy {strategy!r}")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    n = sh.shape[2]
    result = np.zeros(1 + len(pid))
    Jp = np.zeros(n) if want_J_point else None
    h = L.MpHistory()
    h.n, h.ld, h.nsteps, h.strain_comps = n, max(n, 1), sh.shape[0] - 1, sh.shape[1]
    h.strain, h.data = sh.data_ptr(), dh.data_ptr()
    w = np.asarray(weight, dtype=np.float64).reshape(9)
    for k in range(9):
        h.weight[k] = float(w[k])
    h.result = result.ctypes.data
    h.J_point = Jp.ctypes.data if Jp is not None else None
    nw = newton.to_struct()
    rc = lib.cmadx_mp_objective_host(C.byref(material), C.byref(nw), pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                                     C.byref(h), 1 if strategy == "adjoint" else 0, int(device), int(chunk_points))
    L.check(rc, "cmadx_mp_objective_host")
    return (result, Jp) if want_J_point else result


def model_partials(material: L.Material, active_pid, xi: torch.Tensor, xi_prev: torch.Tensor, strain: torch.Tensor,
                   stream: torch.cuda.Stream | None = None) -> dict:
    """The raw AD products of the reference's ``Model`` at given states (``model.py:121-160``):
    ``dC_deps (42, n)``, ``dsig_dxi (42, n)``, ``dsig_deps (36, n)`` and ``dsig_dp (6 P_a, n)`` - see
    ``cmadx_mp_model_partials``.  ``dC/dU_prev`` and ``dcauchy/dxi_prev`` are identically zero."""
    n = xi.shape[1]
    _check_in("xi", xi, 7, n, "cuda"); _check_in("xi_prev", xi_prev, 7, n, "cuda")
    if strain.shape[0] not in (6, 9):
        raise ValueError("strain must have 6 or 9 components")
    _check_in("strain", strain, strain.shape[0], n, "cuda")
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    mk = lambda rows: torch.empty((rows, n), dtype=torch.float64, device=xi.device)  # noqa: E731
    out = {"dC_deps": mk(42), "dsig_dxi": mk(42), "dsig_deps": mk(36)}
    if len(pid):
        out["dsig_dp"] = mk(6 * len(pid))
    p = L.MpPartials()
    p.n, p.ld, p.strain_comps = n, n, strain.shape[0]
    xi, xi_prev, strain = xi.contiguous(), xi_prev.contiguous(), strain.contiguous()
    p.xi, p.xi_prev, p.strain = xi.data_ptr(), xi_prev.data_ptr(), strain.data_ptr()
    for k, v in out.items():
        setattr(p, k, v.data_ptr())
    s = stream if stream is not None else torch.cuda.current_stream(xi.device)
    with torch.cuda.device(xi.device):
        rc = L.lib().cmadx_mp_model_partials(C.byref(material), pid.ctypes.data_as(C.POINTER(C.c_int32)), len(pid),
                                             C.byref(p), s.cuda_stream)
    L.check(rc, "cmadx_mp_model_partials")
    return out


def sym3_eigh(A6: torch.Tensor, vectors: bool = True, stream: torch.cuda.Stream | None = None):
    """Eigen-decomposition of a batch of symmetric 3x3 tensors on the GPU: the drop-in for
    ``sorted_eigen_decomposition`` (cmad/util/jax_eigen_decomposition.py:167-171).  ``A6`` is
    ``(6, n)`` component-major (xx, xy, xz, yy, yz, zz).  Returns ``(w (3, n) ascending, V (3, 3, n))``
    with ``V[:, k, i]`` the k-th eigenvector of tensor i (``None`` when ``vectors`` is false)."""
    A6 = A6.contiguous()
    _check_in("A6", A6, 6, A6.shape[1], "cuda")
    n = A6.shape[1]
    w = torch.empty((3, n), dtype=torch.float64, device=A6.device)
    V = torch.empty((9, n), dtype=torch.float64, device=A6.device) if vectors else None
    s = stream if stream is not None else torch.cuda.current_stream(A6.device)
    with torch.cuda.device(A6.device):
        rc = L.lib().cmadx_sym3_eigh(n, n, A6.data_ptr(), w.data_ptr(), V.data_ptr() if vectors else None, s.cuda_stream)
    L.check(rc, "cmadx_sym3_eigh")
    return w, (V.view(3, 3, n) if vectors else None)


def fp64_peak_tflops(iters: int = 20000) -> float:
    v = C.c_double(0.0)
    s = torch.cuda.current_stream()
    L.check(L.lib().cmadx_fp64_peak(int(iters), C.byref(v), C.c_void_p(s.cuda_stream)), "cmadx_fp64_peak")
    return float(v.value)


def launch_count() -> int:
    return int(L.lib().cmadx_launch_count())


def debug_bail_count(stream: torch.cuda.Stream | None = None) -> int:
    """Points the last J2 radial-return launch handed to the generic kernel."""
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(L.lib().cmadx_debug_bail_count(C.c_void_p(s.cuda_stream)))
