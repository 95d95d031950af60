"""Parameter pytree -> C-ABI material description.

Maps the reference's nested parameter dict (``cmad/parameters/parameters.py``,
leaf layout as in ``tests/support/test_problems.py:9-113`` and the deck
``materials`` sections, e.g. ``examples/elastic_plastic_uniaxial.yaml:37-50``)
to ``cmadx_material_t`` and maps *flat active indices* (sorted-key order) to the
canonical parameter ids whose dC/dp columns the kernels produce.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .parameters import Parameters, tree_leaves_with_path

_YIELD = {"J2": L.YIELD_J2, "hill": L.YIELD_HILL, "hosford": L.YIELD_HOSFORD, "barlat": L.YIELD_BARLAT}
# Yld2004-18p coefficient names in the order of cmadx_material_t::barlat (effective_stress.py:55-78)
BARLAT_KEYS = tuple(f"{p}_{ij}" for p in ("sp", "dp") for ij in ("12", "13", "21", "23", "31", "32", "44", "55", "66"))


@dataclass
class NewtonSettings:
    """Local-Newton settings.  ``traced`` mirrors ``make_newton_solve`` kwargs
    (nonlinear_solver.py:88-100) plus ``DEFAULT_LINE_SEARCH_SETTINGS``
    (line_search.py:40-46); ``imperative`` mirrors ``newton_solve(model)``."""
    mode: str = "traced"
    max_iters: int = 10
    abs_tol: float = 1e-14
    rel_tol: float = 1e-14
    ls_max_evals: int = 4           # traced flavour: probes of the quadratic line search
    ls_sufficient_decrease: float = 1.0e-4
    ls_min_backtrack: float = 0.5
    ls_max_backtrack: float = 0.9
    force_generic: bool = False     # bypass the J2 radial-return specialisation (A/B testing)
    one_pass: bool = False          # reduced Hosford batches: 128-thread blocks without the block-wide Newton vote (A/B)
    stream: bool = False            # material-point batches: the streaming (lane-refill) kernel instead of the
                                    # one-pass generic kernels - A/B testing, identical results
    queue: bool = False             # material-point batches: generic kernel with warp-level parking (mp_update_queue.cu)
    cta: bool = False               # material-point batches: generic kernel with block-level hand-off (mp_update_cta.cu)
    defer_after: int | None = None  # generic kernels' two-pass scheme: None = library default (K = 2 for Hosford a > 8, else off),
                                    # 0 = single pass, K = defer points needing more than K updates
    max_ls_evals: int = 0           # imperative flavour: newton_solve's legacy line search (0 = none, the default)

    def to_struct(self) -> L.Newton:
        if self.mode not in ("traced", "imperative"):
            raise ValueError(f"unknown newton mode {self.mode!r}")
        return L.Newton(L.NEWTON_TRACED if self.mode == "traced" else L.NEWTON_IMPERATIVE,
                        int(self.max_iters), int(self.ls_max_evals if self.mode == "traced" else self.max_ls_evals),
                        (L.NEWTON_F_GENERIC if self.force_generic else 0) | (L.NEWTON_F_ONE_PASS if self.one_pass else 0) | (L.NEWTON_F_STREAM if self.stream else 0) | (L.NEWTON_F_QUEUE if self.queue else 0) | (L.NEWTON_F_CTA if self.cta else 0)
                        | ((0 if self.defer_after is None else (255 if self.defer_after == 0 else
                                                               min(int(self.defer_after), 254))) << 8),
                        float(self.abs_tol), float(self.rel_tol),
                        float(self.ls_sufficient_decrease), float(self.ls_min_backtrack),
                        float(self.ls_max_backtrack))

    @classmethod
    def from_reference_kwargs(cls, max_iters=10, abs_tol=1e-14, rel_tol=1e-14,
                              line_search_settings=None, **_ignored) -> "NewtonSettings":
        ls = {"max evals": 4, "sufficient decrease": 1e-4, "min backtrack factor": 0.5,
              "max backtrack factor": 0.9, **(line_search_settings or {})}
        return cls(mode="traced", max_iters=max_iters, abs_tol=abs_tol, rel_tol=rel_tol,
                   ls_max_evals=ls["max evals"], ls_sufficient_decrease=ls["sufficient decrease"],
                   ls_min_backtrack=ls["min backtrack factor"], ls_max_backtrack=ls["max backtrack factor"])


def _pid_of(path: tuple, k: int, pair: tuple) -> int | None:
    head = path[0]
    if head == "elastic":
        return L.P_EL0 if path[1] == pair[0] else L.P_EL1
    if head == "rotation matrix":
        return L.P_Q00 + k
    if head == "plastic":
        if path[1] == "flow stress":
            if path[2] == "initial yield" and path[3] == "Y":
                return L.P_Y
            if path[2] == "hardening":
                if path[3] == "voce":
                    return {"S": L.P_VOCE_S, "D": L.P_VOCE_D}[path[4]]
                if path[3] == "linear" and path[4] == "K":
                    return L.P_LIN_K
        if path[1] == "effective stress":
            if path[2] == "hill":
                return L.P_HILL_F + "FGHLMN".index(path[3])
            if path[2] == "hosford" and path[3] == "a":
                return L.P_HOSFORD_A
            if path[2] == "barlat":
                return L.P_BARLAT_A if path[3] == "a" else L.P_BARLAT_C0 + BARLAT_KEYS.index(path[3])
            if path[2] == "J2":
                return None          # a leaf with no effect on the model ({"J2": 0.})
    raise ValueError(f"parameter leaf {'/'.join(map(str, path))} is not known to the B200 kernels")


def material_from_values(values: dict, model: str = "small_elastic_plastic",
                         yield_tol: float = 1e-14) -> L.Material:
    """Build ``cmadx_material_t`` from a parameter ``values`` pytree."""
    m = L.Material()
    m.model = {"small_elastic_plastic": L.MODEL_SMALL_ELASTIC_PLASTIC,
               "elastic": L.MODEL_ELASTIC,
               "small_rate_elastic_plastic": L.MODEL_SMALL_RATE_ELASTIC_PLASTIC}[model]
    el = values["elastic"]
    pair = tuple(sorted(el))
    if pair not in L.ELASTIC_PAIRS:
        raise ValueError(f"ElasticConstants needs exactly two of (E, nu, mu, kappa, lambda); got {pair}")
    m.elastic_pair = L.ELASTIC_PAIRS.index(pair)
    m.elastic[0], m.elastic[1] = float(el[pair[0]]), float(el[pair[1]])
    Q = np.asarray(values.get("rotation matrix", np.eye(3)), dtype=np.float64).reshape(9)
    for i in range(9):
        m.Q[i] = float(Q[i])
    m.yield_tol = float(yield_tol)
    if model in ("small_elastic_plastic", "small_rate_elastic_plastic"):
        pl = values["plastic"]
        kind = next(iter(pl["effective stress"]))        # small_elastic_plastic.py:193-198
        if kind not in _YIELD:
            raise NotImplementedError(f"effective stress {kind!r} is outside the B200 path")
        m.yield_ = _YIELD[kind]
        if kind == "hill":
            for i, k in enumerate("FGHLMN"):
                m.hill[i] = float(pl["effective stress"]["hill"][k])
        elif kind == "hosford":
            m.hosford_a = float(pl["effective stress"]["hosford"]["a"])
        elif kind == "barlat":
            for i, k in enumerate(BARLAT_KEYS):
                m.barlat[i] = float(pl["effective stress"]["barlat"][k])
            m.barlat_a = float(pl["effective stress"]["barlat"]["a"])
        m.Y = float(pl["flow stress"]["initial yield"]["Y"])
        mask = 0
        for htype, hp in pl["flow stress"]["hardening"].items():
            if htype == "voce":
                mask |= L.HARD_VOCE
                m.voce_S, m.voce_D = float(hp["S"]), float(hp["D"])
            elif htype == "linear":
                mask |= L.HARD_LINEAR
                m.linear_K = float(hp["K"])
            else:
                raise NotImplementedError(f"hardening {htype!r} is outside the B200 path")
        m.hardening_mask = mask
    return m


def active_param_ids(parameters: Parameters) -> np.ndarray:
    """Canonical parameter id of every active flat entry, in the reference's
    column order (parameters.py:368-377)."""
    pair = tuple(sorted(parameters.values["elastic"]))
    flat = parameters.flat_paths()
    out = []
    for i in parameters.active_idx:
        path, k = flat[int(i)]
        pid = _pid_of(path, k, pair)
        if pid is None:
            raise ValueError(f"active parameter {'/'.join(map(str, path))} does not enter the model")
        out.append(pid)
    return np.asarray(out, dtype=np.int32)


def lame(material: L.Material) -> np.ndarray:
    """[lambda, mu, dlam/de0, dlam/de1, dmu/de0, dmu/de1] via the library."""
    import ctypes as C
    out = (C.c_double * 6)()
    L.check(L.lib().cmadx_lame(C.byref(material), out), "cmadx_lame")
    return np.array(out[:])
