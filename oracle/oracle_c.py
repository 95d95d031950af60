"""ctypes front end of ``oracle_c.cpp`` (the fast batched CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs, never by the
product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")

YIELD = {"J2": 0, "hill": 1, "hosford": 2}
MODEL = {"small_elastic_plastic": 0, "elastic": 1}
# sorted-key pairs ("E" < "kappa" < "lambda" < "mu" < "nu")
ELASTIC_PAIRS = [("E", "nu"), ("E", "mu"), ("E", "kappa"), ("E", "lambda"), ("kappa", "mu"),
                 ("kappa", "nu"), ("kappa", "lambda"), ("lambda", "mu"), ("lambda", "nu"),
                 ("mu", "nu")]
PID = {"EL0": 0, "EL1": 1, "Y": 2, "VOCE_S": 3, "VOCE_D": 4, "LIN_K": 5,
       "HILL_F": 6, "HILL_G": 7, "HILL_H": 8, "HILL_L": 9, "HILL_M": 10, "HILL_N": 11,
       "HOSFORD_A": 12, "Q00": 13}
NUM_PID = 22


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_c.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle_c.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_mp_update.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


@dataclass
class OracleProblem:
    """Flat description of (model, material, solver) for ``oracle_mp_update``."""
    mat: np.ndarray
    cfg: np.ndarray
    sol: np.ndarray
    active_pid: np.ndarray

    @property
    def n_xi(self) -> int:
        return 7 if self.cfg[0] == 0 else 6


def _leaf_path_to_pid(path: tuple, pair: tuple) -> int | None:
    """Map a parameter-pytree leaf path to the canonical parameter id."""
    if path[0] == "elastic":
        return PID["EL0"] if path[1] == pair[0] else PID["EL1"]
    if path[0] == "rotation matrix":
        return PID["Q00"]
    if path[0] == "plastic":
        if path[1] == "flow stress":
            if path[2] == "initial yield":
                return PID["Y"]
            if path[3] == "voce":
                return PID["VOCE_" + path[4]]
            if path[3] == "linear":
                return PID["LIN_K"]
        if path[1] == "effective stress":
            if path[2] == "hill":
                return PID["HILL_" + path[3]]
            if path[2] == "hosford":
                return PID["HOSFORD_A"]
            return None                        # {"J2": 0.}: a leaf with no effect
    raise KeyError(path)


def describe(values: dict, active_idx=None, model: str = "small_elastic_plastic",
             newton_mode: str = "traced", max_iters: int = 10, abs_tol: float = 1e-14,
             rel_tol: float = 1e-14, ls_max_evals: int = 4, max_ls_evals: int = 0, c1: float = 1e-4,
             bmin: float = 0.5, bmax: float = 0.9, strain_comps: int = 6,
             yield_tol: float = 1e-14) -> OracleProblem:
    """Build the flat problem description from a reference-style parameter pytree
    (``values``) and the flat active indices (sorted-key flatten order)."""
    from .cmad_oracle import flatten_tree, leaf_size
    mat = np.zeros(NUM_PID + 1)
    mat[PID["Q00"]:PID["Q00"] + 9] = np.eye(3).reshape(-1)
    el = values["elastic"]
    pair = tuple(sorted(el.keys()))
    cfg = np.zeros(8, dtype=np.int32)
    cfg[0] = MODEL[model]
    cfg[2] = ELASTIC_PAIRS.index(pair)
    mat[0], mat[1] = float(el[pair[0]]), float(el[pair[1]])
    if "rotation matrix" in values:
        mat[PID["Q00"]:PID["Q00"] + 9] = np.asarray(values["rotation matrix"], float).reshape(-1)
    if model == "small_elastic_plastic":
        pl = values["plastic"]
        kind = next(iter(pl["effective stress"]))
        cfg[1] = YIELD[kind]
        if kind == "hill":
            for i, k in enumerate("FGHLMN"):
                mat[PID["HILL_F"] + i] = float(pl["effective stress"]["hill"][k])
        elif kind == "hosford":
            mat[PID["HOSFORD_A"]] = float(pl["effective stress"]["hosford"]["a"])
        mat[PID["Y"]] = float(pl["flow stress"]["initial yield"]["Y"])
        hd = pl["flow stress"]["hardening"]
        if "voce" in hd:
            cfg[3] |= 1
            mat[PID["VOCE_S"]], mat[PID["VOCE_D"]] = float(hd["voce"]["S"]), float(hd["voce"]["D"])
        if "linear" in hd:
            cfg[3] |= 2
            mat[PID["LIN_K"]] = float(hd["linear"]["K"])
    mat[NUM_PID] = yield_tol
    cfg[4] = {"traced": 0, "imperative": 1}[newton_mode]
    # traced: probes of the quadratic line search; imperative: newton_solve's legacy max_ls_evals (0 = none)
    cfg[5], cfg[6], cfg[7] = max_iters, (ls_max_evals if newton_mode == "traced" else max_ls_evals), strain_comps
    sol = np.array([abs_tol, rel_tol, c1, bmin, bmax])
    # flat index -> pid
    pids = []
    for path, v in flatten_tree(values):
        base = _leaf_path_to_pid(path, pair)
        for k in range(leaf_size(v)):
            pids.append(None if base is None else base + (k if path[0] == "rotation matrix" else 0))
    act = []
    for i in (active_idx if active_idx is not None else []):
        if pids[i] is None:
            raise ValueError(f"active parameter {i} has no effect on the model")
        act.append(pids[i])
    return OracleProblem(mat, cfg, sol, np.asarray(act, dtype=np.int32))


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def mp_update(prob: OracleProblem, xi_prev: np.ndarray, strain: np.ndarray,
              want=("xi", "sigma", "dsig_deps", "dC_dp", "iters", "flags", "cnorm"),
              nthreads: int = 0, xi_init: np.ndarray | None = None) -> dict:
    """Batched update on component-major arrays: ``xi_prev (n_xi, N)``,
    ``strain (6|9, N)``.  Returns a dict of the requested outputs."""
    n_xi = prob.n_xi
    xi_prev = np.ascontiguousarray(xi_prev, dtype=np.float64)
    strain = np.ascontiguousarray(strain, dtype=np.float64)
    if xi_init is not None:
        xi_init = np.ascontiguousarray(xi_init, dtype=np.float64)
        assert xi_init.shape == xi_prev.shape
    N = xi_prev.shape[1]
    assert xi_prev.shape == (n_xi, N) and strain.shape == (int(prob.cfg[7]), N)
    na = len(prob.active_pid)
    shapes = {"xi": (n_xi, N), "sigma": (6, N), "dsig_deps": (36, N), "dxi_deps": (n_xi * 6, N),
              "dC_dp": (n_xi * na, N), "dC_dxi": (n_xi * n_xi, N), "dC_dxi_prev": (n_xi * n_xi, N),
              "dsig_dxi": (6 * n_xi, N), "dsig_dp": (6 * na, N)}
    out = {k: (np.zeros(shapes[k]) if k in want else None) for k in shapes}
    ints = {k: (np.zeros(N, dtype=np.int32) if k in want else None) for k in ("iters", "flags", "ls_evals")}
    cnorm = np.zeros(N) if "cnorm" in want else None
    d, i32 = ctypes.c_double, ctypes.c_int
    rc = lib().oracle_mp_update(
        _p(prob.mat, d), _p(prob.cfg, i32), _p(prob.sol, d), _p(prob.active_pid, i32),
        ctypes.c_int(na), ctypes.c_int64(N), ctypes.c_int64(N),
        _p(xi_prev, d), _p(strain, d), _p(xi_init, d),
        _p(out["xi"], d), _p(out["sigma"], d), _p(out["dsig_deps"], d), _p(out["dxi_deps"], d),
        _p(out["dC_dp"], d), _p(out["dC_dxi"], d), _p(out["dC_dxi_prev"], d),
        _p(ints["iters"], i32), _p(ints["flags"], i32), _p(cnorm, d), _p(ints["ls_evals"], i32),
        ctypes.c_int(nthreads), _p(out["dsig_dxi"], d), _p(out["dsig_dp"], d))
    if rc != 0:
        raise RuntimeError(f"oracle_mp_update failed with code {rc}")
    res = {k: v for k, v in out.items() if v is not None}
    res.update({k: v for k, v in ints.items() if v is not None})
    if cnorm is not None:
        res["cnorm"] = cnorm
    return res


def lame(pair_idx: int, p: float, q: float) -> np.ndarray:
    out = np.zeros(6)
    lib().oracle_lame(ctypes.c_int(pair_idx), ctypes.c_double(p), ctypes.c_double(q), _p(out, ctypes.c_double))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
