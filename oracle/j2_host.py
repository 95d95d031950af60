"""ctypes front end of ``j2_host.cpp``: the J2 radial-return routine of the CUDA kernel compiled
for the host (CPU baseline ``kind: "port-handderived"``).

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by ``tests/`` and by
``bench.py``'s CPU-baseline leg, never by the product package.  It compiles product SOURCE
(``cmad_b200/csrc/mp_update_j2_point.cuh``) with the host compiler; the constants the routine
needs come from the product library's ``cmadx_debug_device_structs``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "libj2_host.so")
_SRC = os.path.join(_HERE, "j2_host.cpp")
_CSRC = os.path.join(_ROOT, "cmad_b200", "csrc")


def _cuda_include() -> str:
    for d in (os.environ.get("CUDA_HOME", ""), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    raise RuntimeError("cuda_runtime.h not found (needed for the typedefs the kernel headers use)")


def build(force: bool = False) -> str:
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("mp_update_j2_point.cuh", "j2_radial.cuh", "point_solver.cuh",
                                                      "mp_outputs.cuh", "mp_update.cuh")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fopenmp", "-fPIC", "-w",
                               "-I" + _cuda_include(), "-I" + os.path.join(_ROOT, "include"), "-I" + _CSRC,
                               "-shared", "-o", _SO, _SRC])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.j2_host_mp_update.restype = C.c_int64
    return _lib


def _device_structs(material, newton_struct):
    from cmad_b200 import _lib as L
    sizes = (C.c_int64 * 2)()
    L.check(L.lib().cmadx_debug_device_structs(C.byref(material), C.byref(newton_struct), None, None, sizes),
            "cmadx_debug_device_structs")
    mine = (C.c_int64 * 2)()
    lib().j2_host_struct_sizes(mine)
    if list(sizes) != list(mine):
        raise RuntimeError(f"DevMat / DevNewton layouts differ between nvcc and g++ builds: {list(sizes)} vs {list(mine)}")
    dm, dn = C.create_string_buffer(sizes[0]), C.create_string_buffer(sizes[1])
    L.check(L.lib().cmadx_debug_device_structs(C.byref(material), C.byref(newton_struct), dm, dn, sizes),
            "cmadx_debug_device_structs")
    return dm, dn


def mp_update(material, newton, active_pid, xi_prev: np.ndarray, strain: np.ndarray,
              outputs=("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags"), nthreads: int = 0, out=None):
    """Same contract as ``cmad_b200.mp.mp_update`` on NumPy arrays (component-major).  Returns
    (outputs dict, number of bailed points)."""
    from cmad_b200 import _lib as L
    n = xi_prev.shape[1]
    pid = np.ascontiguousarray(active_pid, dtype=np.int32)
    na = len(pid)
    rows = {"xi": 7, "sigma": 6, "dsig_deps": 36, "dxi_deps": 42, "dC_dp": 7 * na, "dC_dxi": 49, "dC_dxi_prev": 49, "C": 7}
    if out is None:
        out = {k: (np.empty((rows[k], n)) if k in rows else
                   np.empty(n, dtype=np.int32 if k in ("iters", "flags") else np.float64)) for k in outputs}
    b = L.MpBuffers()
    b.n, b.ld, b.strain_comps, b.def_type = n, n, strain.shape[0], 0
    xi_prev, strain = np.ascontiguousarray(xi_prev), np.ascontiguousarray(strain)
    b.xi_prev, b.strain = xi_prev.ctypes.data, strain.ctypes.data
    for k, v in out.items():
        setattr(b, k, v.ctypes.data)
    dm, dn = _device_structs(material, newton.to_struct())
    bails = lib().j2_host_mp_update(dm, dn, pid.ctypes.data_as(C.POINTER(C.c_int32)), na, C.byref(b), int(nthreads))
    return out, int(bails)


def bench_rate(values, params, seed, ts, history_steps, nthreads, sample_points: int = (1 << 22) + 40):
    """updates/s of the host build on the first ``sample_points`` points of the bench workload
    walked through ``ts`` (state carried), all outputs written - the cpu_baseline_handderived
    entry of bench.py.  The sample size is deliberately NOT a power of two: the component-major
    rows are ``sample_points`` doubles apart, and a power-of-two row stride maps the 85 output
    streams of a point onto one cache set (measured 6.6x slower on 8 cores)."""
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, synthetic
    mat, pid, nw = material_from_values(values), active_param_ids(params), NewtonSettings()
    d, d2, a = synthetic.path_params(seed, 0, sample_points)
    xi = np.zeros((7, sample_points))
    out = None
    total, updates, bails = 0.0, 0, 0
    _, _ = mp_update(mat, nw, pid, xi[:, :4096].copy(), synthetic.strain_at_step(d, d2, a, 50)[:, :4096].copy(),
                     nthreads=nthreads)                                     # warm-up (thread pool, page faults)
    bufs = [None, None]
    for j, t in enumerate(ts):
        e = synthetic.strain_at_step(d, d2, a, t)
        t0 = time.perf_counter()
        out, nb = mp_update(mat, nw, pid, xi, e, nthreads=nthreads, out=bufs[j % 2])
        total += time.perf_counter() - t0
        bufs[j % 2] = out
        xi = out["xi"]
        updates += sample_points
        bails += nb
        if total > 20.0:
            break
    return {"value": updates / total, "unit": "updates/s", "cores": int(nthreads) or int(lib().j2_host_max_threads()),
            "kind": "port-handderived",
            "sample": f"first {sample_points} points of the workload, {updates // sample_points} of the timed load steps "
                      f"(state carried), all outputs", "bailed_points": bails,
            "what": "cmad_b200/csrc/mp_update_j2_point.cuh (the CUDA kernel's per-point routine) compiled with "
                    "g++ -O3 -march=native -fopenmp, one point per loop iteration",
            "write_gbs": 680 * updates / total / 1e9}
