"""ctypes binding of ``libcmad_b200.so`` (the C-ABI in ``include/cmad_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc,
sm_100a).  There is no CPU fallback: if the library is missing or fails to
load, importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libcmad_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

SOURCES = ["api.cu", "mp_update.cu", "mp_update_stream.cu", "mp_update_queue.cu", "mp_update_cta.cu", "mp_update_j2.cu", "mp_update_dt.cu", "mp_update_rate.cu", "elastic_update.cu", "mp_sens.cu", "mp_sens_dt.cu", "mp_sens_rate.cu", "mp_hess.cu", "mp_history.cu",
           "fe_block.cu", "fe_tet4.cu", "fe_hex8.cu", "fe_generic.cu", "fe_tet4x4.cu", "fe_rate.cu", "fe_mixed.cu", "fe_post.cu", "fe_vjp.cu", "fe_scatter.cu", "sym3_eigh.cu", "mp_partials.cu", "mp_update_rate_dt.cu", "mp_sens_rate_dt.cu"]

# ---- enums (mirror include/cmad_b200.h) ---------------------------------
OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = range(5)
VERSION = 100        # CMADX_VERSION of include/cmad_b200.h this binding was written against
MODEL_SMALL_ELASTIC_PLASTIC, MODEL_ELASTIC, MODEL_SMALL_RATE_ELASTIC_PLASTIC = 0, 1, 2
QOI_CALIBRATION, QOI_UNIAXIAL_CALIBRATION = 0, 1
YIELD_J2, YIELD_HILL, YIELD_HOSFORD, YIELD_BARLAT = 0, 1, 2, 3
DEF_FULL_3D, DEF_PLANE_STRESS, DEF_UNIAXIAL_STRESS = 0, 1, 2
ELASTIC_PAIRS = [("E", "nu"), ("E", "mu"), ("E", "kappa"), ("E", "lambda"), ("kappa", "mu"),
                 ("kappa", "nu"), ("kappa", "lambda"), ("lambda", "mu"), ("lambda", "nu"),
                 ("mu", "nu")]
HARD_VOCE, HARD_LINEAR = 1, 2
(P_EL0, P_EL1, P_Y, P_VOCE_S, P_VOCE_D, P_LIN_K, P_HILL_F, P_HILL_G, P_HILL_H, P_HILL_L,
 P_HILL_M, P_HILL_N, P_HOSFORD_A, P_Q00) = range(14)
P_BARLAT_C0 = P_Q00 + 9
P_BARLAT_A = P_BARLAT_C0 + 18
NUM_PARAM_IDS = P_BARLAT_A + 1
MAX_ACTIVE = 16
NEWTON_TRACED, NEWTON_IMPERATIVE = 0, 1
NEWTON_F_GENERIC = 1
NEWTON_F_ONE_PASS = 2
NEWTON_F_STREAM = 4
NEWTON_F_QUEUE = 8
NEWTON_F_CTA = 16


class Material(C.Structure):
    _fields_ = [("model", C.c_int32), ("yield_", C.c_int32), ("elastic_pair", C.c_int32),
                ("hardening_mask", C.c_int32), ("elastic", C.c_double * 2), ("Y", C.c_double),
                ("voce_S", C.c_double), ("voce_D", C.c_double), ("linear_K", C.c_double),
                ("hill", C.c_double * 6), ("hosford_a", C.c_double), ("Q", C.c_double * 9),
                ("yield_tol", C.c_double), ("barlat", C.c_double * 18), ("barlat_a", C.c_double)]


class MpPartials(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64), ("strain_comps", C.c_int32), ("reserved", C.c_int32),
                ("xi", C.c_void_p), ("xi_prev", C.c_void_p), ("strain", C.c_void_p), ("dC_deps", C.c_void_p),
                ("dsig_dxi", C.c_void_p), ("dsig_deps", C.c_void_p), ("dsig_dp", C.c_void_p)]


class Newton(C.Structure):
    _fields_ = [("mode", C.c_int32), ("max_iters", C.c_int32), ("ls_max_evals", C.c_int32),
                ("flags", C.c_int32), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("ls_c1", C.c_double), ("ls_bmin", C.c_double), ("ls_bmax", C.c_double)]


class MpBuffers(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64), ("strain_comps", C.c_int32),
                ("def_type", C.c_int32), ("xi_prev", C.c_void_p), ("strain", C.c_void_p), ("xi_init", C.c_void_p),
                ("xi", C.c_void_p), ("sigma", C.c_void_p), ("dsig_deps", C.c_void_p),
                ("dxi_deps", C.c_void_p), ("dC_dp", C.c_void_p), ("dC_dxi", C.c_void_p),
                ("dC_dxi_prev", C.c_void_p), ("iters", C.c_void_p), ("flags", C.c_void_p),
                ("cnorm", C.c_void_p), ("C", C.c_void_p), ("strain_prev", C.c_void_p)]


class MpHistory(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64), ("nsteps", C.c_int32),
                ("strain_comps", C.c_int32), ("strain", C.c_void_p), ("data", C.c_void_p),
                ("weight", C.c_double * 9), ("xi_hist", C.c_void_p), ("iters_hist", C.c_void_p),
                ("result", C.c_void_p), ("workspace", C.c_void_p), ("J_point", C.c_void_p),
                ("qoi_kind", C.c_int32), ("reserved_", C.c_int32), ("weight_steps", C.c_void_p)]


class FeBlock(C.Structure):
    _fields_ = [("n_elems", C.c_int64), ("n_dofs", C.c_int64), ("n_basis", C.c_int32),
                ("n_ip", C.c_int32), ("elem_eq", C.c_void_p), ("U", C.c_void_p),
                ("xi_prev", C.c_void_p), ("grad_N", C.c_void_p), ("det", C.c_void_p),
                ("quad_w", C.c_void_p), ("xi", C.c_void_p), ("R_elem", C.c_void_p),
                ("K_elem", C.c_void_p), ("R_global", C.c_void_p), ("sigma", C.c_void_p),
                ("iters", C.c_void_p), ("flags", C.c_void_p), ("U_prev", C.c_void_p)]


class FeMixed(C.Structure):
    _fields_ = [("elem_eq_p", C.c_void_p), ("N", C.c_void_p), ("h", C.c_void_p),
                ("stab_mult", C.c_double), ("R_p_elem", C.c_void_p), ("K_up", C.c_void_p),
                ("K_pu", C.c_void_p), ("K_pp", C.c_void_p), ("R_global", C.c_void_p)]


class CmadxError(RuntimeError):
    pass


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + INCLUDE]
OBJ_DIR = os.path.join(_HERE, "lib", "obj")


def nvcc_command(out: str = LIB_PATH) -> list[str]:
    """Single-shot equivalent of what :func:`build` does (compile all + link)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    return ["nvcc"] + NVCC_FLAGS + ["-shared", "-o", out] + srcs


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (``nvcc -gencode
    arch=compute_100a,code=sm_100a -lineinfo``) into
    ``cmad_b200/lib/libcmad_b200.so``; objects are compiled in parallel and only
    when their source (or any header) is newer."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "cmad_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(subprocess.check_call, jobs))
    if jobs or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                               "-o", LIB_PATH] + objs)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raise loudly if it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmadxError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  cmad_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.cmadx_version.restype = C.c_int
    L.cmadx_error_string.restype = C.c_char_p
    L.cmadx_error_string.argtypes = [C.c_int]
    L.cmadx_last_cuda_error.restype = C.c_char_p
    L.cmadx_launch_count.restype = C.c_int64
    L.cmadx_debug_bail_count.restype = C.c_int64
    L.cmadx_debug_bail_count.argtypes = [C.c_void_p]
    L.cmadx_lame.argtypes = [C.POINTER(Material), C.POINTER(C.c_double)]
    mp_args = [C.POINTER(Material), C.POINTER(Newton), C.POINTER(C.c_int32), C.c_int32,
               C.POINTER(MpBuffers)]
    L.cmadx_mp_update.argtypes = mp_args + [C.c_void_p]
    L.cmadx_mp_update_host.argtypes = mp_args + [C.c_int, C.c_int64]
    L.cmadx_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_void_p]
    L.cmadx_mp_objective_workspace_bytes.restype = C.c_int64
    L.cmadx_mp_objective_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    L.cmadx_mp_forward_history.argtypes = [C.POINTER(Material), C.POINTER(Newton),
                                           C.POINTER(MpHistory), C.c_void_p]
    obj_args = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32, C.POINTER(MpHistory), C.c_void_p]
    L.cmadx_mp_objective_adjoint.argtypes = obj_args
    L.cmadx_mp_objective_host.argtypes = [C.POINTER(Material), C.POINTER(Newton), C.POINTER(C.c_int32), C.c_int32,
                                          C.POINTER(MpHistory), C.c_int, C.c_int, C.c_int64]
    L.cmadx_mp_objective_direct.argtypes = obj_args
    L.cmadx_mp_objective_hessian.argtypes = obj_args[:-1] + [C.c_int32, C.c_void_p]
    L.cmadx_mp_hessian_workspace_bytes.restype = C.c_int64
    L.cmadx_mp_hessian_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    L.cmadx_fe_block_assemble.argtypes = [C.POINTER(Material), C.POINTER(Newton),
                                          C.POINTER(FeBlock), C.c_void_p]
    L.cmadx_fe_block_assemble_mixed.argtypes = [C.POINTER(Material), C.POINTER(Newton),
                                                C.POINTER(FeBlock), C.POINTER(FeMixed), C.c_void_p]
    L.cmadx_fe_cauchy_at_ips.argtypes = [C.POINTER(Material), C.POINTER(FeBlock), C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    L.cmadx_embedded_plan_create.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64, C.c_int64,
                                             C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p)]
    L.cmadx_embedded_plan_destroy.argtypes = [C.c_void_p]
    L.cmadx_embedded_apply.argtypes = [C.c_void_p] * 8
    L.cmadx_fe_block_jvp.argtypes = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32,
                                     C.POINTER(C.c_double), C.POINTER(FeBlock), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_fe_vjp_workspace_bytes.restype = C.c_int64
    L.cmadx_fe_vjp_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    L.cmadx_fe_block_vjp.argtypes = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32,
                                     C.POINTER(FeBlock), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_fe_block_jvp_mixed.argtypes = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32,
                                           C.POINTER(C.c_double), C.POINTER(FeBlock), C.POINTER(FeMixed),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_fe_block_vjp_mixed.argtypes = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32,
                                           C.POINTER(FeBlock), C.POINTER(FeMixed), C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_fe_block_vjp_disp.argtypes = [C.POINTER(Material), C.POINTER(FeBlock), C.POINTER(FeMixed),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_segment_plan_create.argtypes = [C.POINTER(C.c_int64), C.c_int64, C.c_int64,
                                            C.POINTER(C.c_void_p)]
    L.cmadx_segment_plan_destroy.argtypes = [C.c_void_p]
    L.cmadx_segment_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.cmadx_mp_model_partials.argtypes = [C.POINTER(Material), C.POINTER(C.c_int32), C.c_int32,
                                          C.POINTER(MpPartials), C.c_void_p]
    L.cmadx_sym3_eigh.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_index_gather.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cmadx_index_scatter.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    # a stale library (the .so is git-ignored and rebuilt by mtime) with a drifted struct layout
    # would be driven with mis-laid-out structs: refuse it here instead of corrupting memory
    if L.cmadx_version() != VERSION:
        raise CmadxError(f"{LIB_PATH}: library version {L.cmadx_version()} != binding version {VERSION}; rebuild")
    sizes = (C.c_int64 * 5)()
    L.cmadx_struct_sizes(sizes)
    want = [C.sizeof(t) for t in (Material, Newton, MpBuffers, MpHistory, FeBlock)]
    if list(sizes) != want:
        raise CmadxError(f"{LIB_PATH}: struct sizes {list(sizes)} != ctypes layout {want}; rebuild "
                         "(`python -c 'import __graft_entry__ as g; g.build()'`)")
    _lib = L
    return L


def check(rc: int, what: str = "cmadx call") -> None:
    if rc != OK:
        L = lib()
        msg = L.cmadx_error_string(rc).decode()
        if rc == ECUDA:
            msg += " (" + L.cmadx_last_cuda_error().decode() + ")"
        if rc == EINVAL:
            raise ValueError(f"{what}: {msg}")
        if rc == EUNSUPPORTED:
            raise NotImplementedError(f"{what}: {msg}")
        raise CmadxError(f"{what}: {msg}")


def exported_symbols() -> list[str]:
    """Function names declared in include/cmad_b200.h (for the load test)."""
    import re
    text = open(os.path.join(INCLUDE, "cmad_b200.h")).read()
    return sorted(set(re.findall(r"\b(cmadx_[a-z0-9_]+)\s*\(", text)))
