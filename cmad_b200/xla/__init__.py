"""JAX FFI binding of ``libcmad_b200.so`` (``cmad_b200_xla.cc``): the custom calls a JAX program -
the reference itself - uses to reach the B200 kernels with its own device arrays, on XLA's
stream, x64 mode untouched.

JAX is not installable in the build image, so nothing here is exercised at test time except
:func:`available` (False) and the C++ source's syntax check; the module is the code a
deployment with JAX runs:

    from cmad_b200 import xla
    xla.build()            # g++ against jax.ffi.include_dir(), links libcmad_b200.so
    xla.register()         # jax.ffi.register_ffi_target(..., platform="CUDA")
    R_e, K_e, xi = xla.fe_block(eq, U, xi_prev, grad_N, det, quad_w, material, newton)

``cmad_b200.cmad_plugin.FfiBackend`` routes the reference's ``assemble_element_block`` through
these calls."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .. import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(_HERE, "cmad_b200_xla.cc")
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcmad_b200_xla.so")

# FFI target name -> exported handler symbol
TARGETS = {
    "cmadx_mp_update": "CmadxMpUpdate",
    "cmadx_mp_update_full": "CmadxMpUpdateFull",
    "cmadx_mp_model_partials": "CmadxMpModelPartials",
    "cmadx_sym3_eigh": "CmadxSym3Eigh",
    "cmadx_fe_block": "CmadxFeBlock",
    "cmadx_fe_block_residual": "CmadxFeBlockResidual",
    "cmadx_fe_block_mixed": "CmadxFeBlockMixed",
    "cmadx_fe_block_jvp": "CmadxFeBlockJvp",
    "cmadx_fe_block_vjp": "CmadxFeBlockVjp",
    "cmadx_fe_block_vjp_disp": "CmadxFeBlockVjpDisp",
    "cmadx_fe_block_jvp_mixed": "CmadxFeBlockJvpMixed",
    "cmadx_fe_block_vjp_mixed": "CmadxFeBlockVjpMixed",
}
_registered = False


def available() -> bool:
    """True where JAX with the FFI headers is importable (never in the build image)."""
    try:
        import jax
        return hasattr(jax, "ffi") and os.path.isdir(jax.ffi.include_dir())
    except Exception:
        return False


def build_command(include_dir: str, out: str = LIB_PATH) -> list[str]:
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.dirname(L.LIB_PATH)
    return ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + include_dir, "-I" + L.INCLUDE,
            "-I" + os.path.join(cuda, "include"), SOURCE, "-o", out, "-L" + libdir, "-lcmad_b200",
            "-Wl,-rpath," + libdir]


def build(force: bool = False) -> str:
    """Compile the handlers where ``jax.ffi.include_dir()`` exists; raises where it does not."""
    if not available():
        raise RuntimeError("cmad_b200.xla.build(): JAX (jax.ffi.include_dir()) is not available")
    import jax
    L.build()
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(SOURCE), os.path.getmtime(L.LIB_PATH)):
        subprocess.check_call(build_command(jax.ffi.include_dir()))
    return LIB_PATH


def register() -> None:
    """``jax.ffi.register_ffi_target`` for every handler (platform CUDA)."""
    global _registered
    if _registered:
        return
    import jax
    so = C.CDLL(build())
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(so, symbol)), platform="CUDA")
    _registered = True


def struct_bytes(s: C.Structure) -> np.ndarray:
    """A ctypes mirror of a C-ABI struct as the uint8 array attribute the handlers unpack."""
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def _call(target, shapes, *args, **attrs):
    import jax
    import jax.numpy as jnp
    register()
    out = tuple(jax.ShapeDtypeStruct(tuple(int(x) for x in shp), dt) for shp, dt in shapes)
    del jnp
    return jax.ffi.ffi_call(target, out)(*args, **attrs)


def mp_update(xi_prev, strain, material, newton, active_pid, def_type: int = 0):
    """``(xi, sigma, dsig_deps, dC_dp, iters, flags)`` for ``[comps][n]`` device arrays."""
    import jax.numpy as jnp
    n_xi, n = xi_prev.shape
    na = max(len(active_pid), 1)
    ns = {0: 6, 1: 3, 2: 1}[int(def_type)]
    f, i = jnp.float64, jnp.int32
    return _call("cmadx_mp_update",
                 [((n_xi, n), f), ((6, n), f), ((6 * ns, n), f), ((n_xi * na, n), f), ((n,), i), ((n,), i)],
                 xi_prev, strain, material=struct_bytes(material), newton=struct_bytes(newton),
                 active_pid=np.asarray(active_pid, np.int32), def_type=np.int32(def_type))


def fe_block(elem_eq, U, xi_prev, grad_N, det, quad_w, material, newton, want_K: bool = True):
    """``(R_elem, K_elem, xi)`` (``want_K=False``: ``(R_elem, xi)``) of one COUPLED block."""
    import jax.numpy as jnp
    n_e, n_ip, n_b, _ = grad_N.shape
    nd = 3 * n_b
    f = jnp.float64
    attrs = dict(material=struct_bytes(material), newton=struct_bytes(newton))
    if want_K:
        return _call("cmadx_fe_block", [((n_e, nd), f), ((n_e, nd, nd), f), (xi_prev.shape, f)],
                     elem_eq, U, xi_prev, grad_N, det, quad_w, **attrs)
    return _call("cmadx_fe_block_residual", [((n_e, nd), f), (xi_prev.shape, f)],
                 elem_eq, U, xi_prev, grad_N, det, quad_w, **attrs)


def fe_block_mixed(elem_eq, elem_eq_p, U, xi_prev, grad_N, det, quad_w, N, h, material, newton, stab_mult: float):
    """``(R_u, R_p, K_uu, K_up, K_pu, K_pp, xi)`` of one mixed u-p block."""
    import jax.numpy as jnp
    n_e, n_ip, n_b, _ = grad_N.shape
    nu, npd = 3 * n_b, n_b
    f = jnp.float64
    return _call("cmadx_fe_block_mixed",
                 [((n_e, nu), f), ((n_e, npd), f), ((n_e, nu, nu), f), ((n_e, nu, npd), f), ((n_e, npd, nu), f),
                  ((n_e, npd, npd), f), (xi_prev.shape, f)],
                 elem_eq, elem_eq_p, U, xi_prev, grad_N, det, quad_w, N, h,
                 material=struct_bytes(material), newton=struct_bytes(newton), stab_mult=np.float64(stab_mult))


def fe_block_jvp(elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, dxi_prev, dU, material, active_pid, dp):
    """``(dR_elem, dxi)``: the JVP rule of the block w.r.t. ``(params, xi_prev, U)``."""
    import jax.numpy as jnp
    n_e, n_ip, n_b, _ = grad_N.shape
    f = jnp.float64
    return _call("cmadx_fe_block_jvp", [((n_e, 3 * n_b), f), (xi_prev.shape, f)],
                 elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, dxi_prev, dU,
                 material=struct_bytes(material), active_pid=np.asarray(active_pid, np.int32),
                 dp=np.asarray(dp, np.float64))


def fe_block_vjp(elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, Rbar, xibar, material, active_pid):
    """``(pbar, xibar_prev)``: the transpose of :func:`fe_block_jvp` (what ``jax.grad`` needs)."""
    import jax.numpy as jnp
    n_e, n_ip, n_b, _ = grad_N.shape
    na = len(active_pid)
    ws = int(L.lib().cmadx_fe_vjp_workspace_bytes(C.c_int64(n_e), C.c_int32(n_ip), C.c_int32(na))) // 8
    f = jnp.float64
    pbar, xibar_prev, _ = _call("cmadx_fe_block_vjp", [((na,), f), (xi_prev.shape, f), ((max(ws, 1),), f)],
                                elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, Rbar, xibar,
                                material=struct_bytes(material), active_pid=np.asarray(active_pid, np.int32))
    return pbar, xibar_prev


def fe_block_vjp_disp(elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, Rbar, xibar, material):
    """``(Ubar_ip (n_e * n_ip, 3 n_b), xibar_prev)``: displacement cotangent rows per point."""
    import jax.numpy as jnp
    n_e, n_ip, n_b, _ = grad_N.shape
    f = jnp.float64
    return _call("cmadx_fe_block_vjp_disp", [((n_e * n_ip, 3 * n_b), f), (xi_prev.shape, f)],
                 elem_eq, U, xi_prev, grad_N, det, quad_w, xi_state, Rbar, xibar, material=struct_bytes(material))
