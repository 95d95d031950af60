"""The fast C++ dual-number oracle must agree with the torch.func AD oracle
(values to 1e-10, iteration counts and branch flags exactly), and the line
search must satisfy the reference's own unit tests."""
import numpy as np
import pytest
import torch

from oracle import cmad_oracle as co, oracle_c as oc
from tests.helpers import UP, param_tree, random_strains, rel_err, rotation_matrix

WANT = ("xi", "sigma", "dsig_deps", "dxi_deps", "dC_dp", "dC_dxi", "dC_dxi_prev", "iters", "flags", "cnorm")


def _sym(e):
    return torch.tensor([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]], dtype=co.DT)


@pytest.mark.parametrize("kind,mode,hard,rot", [
    ("J2", "traced", ("voce",), None), ("J2", "imperative", ("voce", "linear"), None),
    ("hill", "traced", ("voce",), None), ("hosford", "traced", ("voce",), None),
    ("J2", "traced", ("linear",), "rot"), ("hill", "imperative", ("voce",), "rot"),
])
def test_c_oracle_matches_torch_oracle(kind, mode, hard, rot):
    rng = np.random.default_rng(7)
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    active = ("E", "nu", "D", "S", "Y", "K") + (tuple("FGHLMN") if kind == "hill" else ())
    Q = rotation_matrix([1, 2, 3], 0.7) if rot else None
    values, act, tr = param_tree(kind, hard, hill=hill, active=active, rotation=Q)
    P = co.OracleParameters(values, act, tr)
    spec = co.ModelSpec()
    kw = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
    prob = oc.describe(values, P.active_idx, newton_mode=mode, **kw)
    N = 6
    e1 = random_strains(rng, N, diag_only=(kind == "hosford"))
    r1 = oc.mp_update(prob, np.zeros((7, N)), e1, want=WANT)
    e2 = 1.4 * e1 + 0.2 * random_strains(rng, N, diag_only=(kind == "hosford"))
    r2 = oc.mp_update(prob, r1["xi"], e2, want=WANT)
    params = co.to_torch_tree(values)
    solver = co.newton_traced if mode == "traced" else co.newton_imperative
    for xprev, e, r in ((np.zeros((7, N)), e1, r1), (r1["xi"], e2, r2)):
        for i in range(N):
            gu = _sym(e[:, i]); xp = torch.as_tensor(xprev[:, i].copy())
            x, info = solver(xp, params, gu, gu, spec, **kw)
            assert info.iters == r["iters"][i]
            assert (info.flag_entry | (info.flag_exit << 1)) == r["flags"][i]
            assert rel_err(r["xi"][:, i], x.numpy()) < 1e-10
            sg = co.sep_cauchy(x, xp, params, gu, gu, spec).numpy()
            assert rel_err(r["sigma"][:, i], [sg[a] for a in UP]) < 1e-10
            A = co.dC_dxi(x, xp, params, gu, gu, spec).numpy()
            assert rel_err(r["dC_dxi"][:, i].reshape(7, 7), A) < 1e-10
            B = co.dC_dxi_prev(x, xp, params, gu, gu, spec).numpy()
            assert rel_err(r["dC_dxi_prev"][:, i].reshape(7, 7), B) < 1e-10
            dp = P.active_params_jacobian(
                co.tree_map(lambda t: t.numpy(), co.dC_dparams(x, xp, params, gu, gu, spec)), 7)
            if np.abs(dp).max() > 0:
                assert rel_err(r["dC_dp"][:, i].reshape(7, -1), dp) < 1e-10
            D = co.consistent_tangent(x, xp, params, gu, gu, spec).numpy()
            D6 = np.array([[D[i1, j1, k, l] + (D[i1, j1, l, k] if k != l else 0.0)
                            for (k, l) in UP] for (i1, j1) in UP])
            assert rel_err(r["dsig_deps"][:, i].reshape(6, 6), D6) < 1e-10
    assert r2["flags"].max() >= 2          # at least one plastic point exercised


def test_c_oracle_elastic_model_matches_torch():
    values = {"elastic": {"kappa": 100.0, "mu": 50.0}}
    prob = oc.describe(values, [0, 1], model="elastic", max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
    rng = np.random.default_rng(3)
    e = random_strains(rng, 5)
    r = oc.mp_update(prob, np.zeros((6, 5)), e, want=WANT)
    params = co.to_torch_tree(values); spec = co.ModelSpec(kind="elastic")
    for i in range(5):
        gu = _sym(e[:, i])
        x, info = co.newton_traced(np.zeros(6), params, gu, gu, spec, max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
        assert info.iters == r["iters"][i] == 1
        assert rel_err(r["xi"][:, i], x.numpy()) < 1e-12
        dp = co.dC_dparams(x, torch.zeros(6, dtype=co.DT), params, gu, gu, spec)
        dpm = np.stack([dp["elastic"]["kappa"].numpy(), dp["elastic"]["mu"].numpy()], axis=1)
        assert rel_err(r["dC_dp"][:, i].reshape(6, 2), dpm) < 1e-10


def test_lame_pairs_consistent():
    """Every supported pair (elastic_constants.py:54-104) maps to the same Lame pair."""
    E, nu = 200e3, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    kappa = lam + 2 * mu / 3
    vals = {"E": E, "nu": nu, "mu": mu, "kappa": kappa, "lambda": lam}
    for idx, (a, b) in enumerate(oc.ELASTIC_PAIRS):
        out = oc.lame(idx, vals[a], vals[b])
        assert abs(out[0] - lam) < 1e-9 * lam and abs(out[1] - mu) < 1e-9 * mu
        lt, mt = co.lame_from_params({a: torch.tensor(vals[a], dtype=co.DT), b: torch.tensor(vals[b], dtype=co.DT)})
        assert abs(float(lt) - out[0]) < 1e-9 * lam and abs(float(mt) - out[1]) < 1e-9 * mu


# ---- line search: the reference's unit tests (tests/util/test_line_search.py:15-163)
def test_quad_min_exact_on_quadratic():
    # phi(a) = (a - 0.3)^2: phi0 = .09, dphi0 = -.6, phi(1) = .49 -> minimiser 0.3
    assert abs(co.quad_min(0.09, -0.6, 1.0, 0.49) - 0.3) < 1e-14
    assert co.quad_min(1.0, -1.0, 1.0, 0.0) == 0.5      # degenerate curvature -> a/2


def test_line_search_accepts_full_step():
    calls = []
    def ev(a):
        calls.append(a); return 0.5 * (1 - a) ** 2, "aux%g" % a
    a, aux, n = co.line_search(ev, 0.5, -1.0, co.DEFAULT_LINE_SEARCH_SETTINGS, "init")
    assert a == 1.0 and n == 1 and aux == "aux1"


def test_line_search_backtracks_and_clips():
    # merit rises steeply at full step: first trial rejected, contraction clipped to [0.5, 0.9] a
    def ev(a):
        return 0.5 * (1 - 4 * a) ** 2, a
    a, aux, n = co.line_search(ev, 0.5, -1.0, co.DEFAULT_LINE_SEARCH_SETTINGS, None)
    assert n >= 2 and a < 1.0 and aux == a
    # non-finite merit halves the step
    seq = []
    def ev2(a):
        seq.append(a); return (float("nan") if a > 0.3 else 0.0), a
    a, aux, n = co.line_search(ev2, 0.5, -1.0, co.DEFAULT_LINE_SEARCH_SETTINGS, None)
    assert seq[:3] == [1.0, 0.5, 0.25] and a == 0.25


def test_line_search_returns_best_when_never_accepted():
    def ev(a):
        return 1.0 + a, a          # always worse than phi0 = 0.5
    a, aux, n = co.line_search(ev, 0.5, -1.0, co.DEFAULT_LINE_SEARCH_SETTINGS, "init")
    assert n == 4 and aux == a and a < 0.2     # lowest merit = smallest step tried
