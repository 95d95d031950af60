"""Parity against golden vectors produced by EXECUTING THE REFERENCE'S OWN,
UNMODIFIED SOURCE for the hot path (tests/golden/make_reference_golden.py runs
/root/reference/cmad/... on the NumPy `jax` stand-in of tests/golden/jaxshim/;
fixtures tests/golden/ref_*.npz, committed; /root/reference is not read here).

What the fixtures hold, all computed by the reference's code:
* ref_traced_newton.npz   - `make_newton_solve(model._residual)` (nonlinear_solver.py:88-174)
  on J2 / Hill / rotated anisotropic Hill / Hosford(a=4, a=100): xi, the Newton
  update count of the loop, plastic flags of `cond_residual`, `model.cauchy`,
  the IFT tangents d(xi, sigma)/d(grad_u) and d(xi, sigma)/dp through the
  `custom_jvp` rule, and `Model`'s AD products dC/dxi, dC/dxi_prev, dC/dp;
* ref_imperative_newton.npz - the `Model` object driven by the imperative
  `newton_solve` (nonlinear_solver.py:14-85): per-step (ii, ||C||), xi, Sigma;
* ref_mp_objectives.npz   - `MPAdjointObjective` / `MPDirectObjective` with the
  `Calibration` QoI: J and the (transformed) gradient.

CPU tests: the C++ oracle (and the torch oracle on a subset) against the
fixtures.  GPU tests: the CUDA path through the C-ABI against the fixtures.
Values 1e-10 relative (tangents / gradients 1e-9), Newton counts and branch
flags exactly equal.
"""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from oracle import mp_objective_np as mo, oracle_c as oc
from tests.golden.materials import NEWTON, active_all_scalars, const_like, material, objective_trees
from tests.helpers import UP, rel_err

G = os.path.join(os.path.dirname(__file__), "golden")
TR = np.load(os.path.join(G, "ref_traced_newton.npz"))
IM = np.load(os.path.join(G, "ref_imperative_newton.npz"))
OB = np.load(os.path.join(G, "ref_mp_objectives.npz"))

TRACED = sorted({k.rsplit(".", 1)[0] for k in TR.files})
IMPER = sorted({k.rsplit(".", 1)[0] for k in IM.files})
OBJ = sorted({k.rsplit(".", 1)[0] for k in OB.files})
# Round 1 compared the notch material (Hosford a = 100, 500 iterations, 100 probes) statistically.
# Both the C++ oracle and the CUDA kernels reproduce the reference's counts and flags on every
# point of that fixture, and the reference's own answers do not move under a one-ulp change of
# its inputs (tests/golden/ref_knife_edge.npz): the comparison is exact for every case now.
LOOSE: set = set()

_ROW9 = [3 * i + j for i, j in UP]                                   # packed component -> row-major 3x3 entry
_COLS = [((3 * k + l,) if k == l else (3 * k + l, 3 * l + k)) for k, l in UP]


def _sym_cols(M9):
    """(..., 9) derivative w.r.t. the row-major entries of grad_u -> (..., 6)
    derivative w.r.t. the symmetric strain components (both entries moved)."""
    return np.stack([sum(M9[..., c] for c in cols) for cols in _COLS], axis=-1)


def _newton_kw(key):
    kw = dict(NEWTON[key])
    ls = kw.pop("line_search_settings", {})
    if "max evals" in ls:
        kw["ls_max_evals"] = ls["max evals"]
    return kw


def _setup(case):
    kind, key = case.split(".")
    values = material(kind)
    P = Parameters(values, active_all_scalars(values), const_like(values, None))
    return kind, key, values, P


def _check_traced(case, run_update):
    """`run_update(values, P, newton_kw, xi_prev (7,n), grad_u (9,n)) -> dict` of
    component-major outputs; compared step by step with the reference's."""
    kind, key, values, P = _setup(case)
    g = {k: TR[f"{case}.{k}"] for k in ("grad_u", "xi_prev", "xi", "iters", "flags", "sigma", "dsig_dgradu",
                                       "dxi_dgradu", "dC_dp", "dC_dxi", "dC_dxi_prev", "dxi_dp", "dsig_dp")}
    assert list(TR[f"{case}.param_names"]) == [f"['{n}']" for n in P._names] or \
        len(TR[f"{case}.param_names"]) == len(P._names)
    aidx = np.asarray(P.active_idx)
    n_plastic = 0
    for s in range(g["xi"].shape[0]):
        out = run_update(values, P, _newton_kw(key), g["xi_prev"][s].T.copy(), g["grad_u"][s].T.copy())
        it, fl = np.asarray(out["iters"]), np.asarray(out["flags"])
        same = (it == g["iters"][s]) & (fl == g["flags"][s])
        if case in LOOSE:
            assert same.mean() >= 0.85, (case, s, it, g["iters"][s])
        else:
            assert same.all(), (case, s, np.flatnonzero(~same), it[~same], g["iters"][s][~same])
        n_plastic += int((g["flags"][s] & 2).astype(bool).sum())
        n = it.size
        ref = {
            "xi": g["xi"][s].T,
            "sigma": g["sigma"][s].T,
            "dsig_deps": _sym_cols(g["dsig_dgradu"][s])[:, _ROW9, :].reshape(n, 36).T,
            "dxi_deps": _sym_cols(g["dxi_dgradu"][s]).reshape(n, 42).T,
            "dC_dxi": g["dC_dxi"][s].reshape(n, 49).T,
            "dC_dxi_prev": g["dC_dxi_prev"][s].reshape(n, 49).T,
            "dC_dp": g["dC_dp"][s][:, :, aidx].reshape(n, -1).T,
        }
        tol = {"xi": 1e-10, "sigma": 1e-10}
        for k, r in ref.items():
            got = np.asarray(out[k])[:, same]
            assert rel_err(got, r[:, same]) < tol.get(k, 1e-9), (case, s, k, rel_err(got, r[:, same]))
        # the reference's IFT parameter tangents equal -A^{-1} dC/dp and the chain rule
        # through cauchy, evaluated from OUR derivative outputs
        A = np.asarray(out["dC_dxi"]).T.reshape(n, 7, 7)
        dCdp = np.asarray(out["dC_dp"]).T.reshape(n, 7, len(aidx))
        dxdp = -np.linalg.solve(A, dCdp)
        ok = same & ((g["flags"][s] & 2) > 0)
        if ok.any():
            assert rel_err(dxdp[ok], g["dxi_dp"][s][:, :, aidx][ok]) < 1e-8, (case, s, "dxi_dp")
            if "dsig_dxi" in out:           # partial stress derivatives: oracle only (K6 uses them on the GPU)
                ds_dx = np.asarray(out["dsig_dxi"]).T.reshape(n, 6, 7)
                ds_dp = np.asarray(out["dsig_dp"]).T.reshape(n, 6, len(aidx))
                tot = ds_dp + ds_dx @ dxdp
                assert rel_err(tot[ok], g["dsig_dp"][s][:, _ROW9][:, :, aidx][ok]) < 1e-8, (case, s, "dsig_dp")
    assert n_plastic > 0


WANT = ("xi", "sigma", "dsig_deps", "dxi_deps", "dC_dp", "dC_dxi", "dC_dxi_prev", "iters", "flags",
        "dsig_dxi", "dsig_dp")


@pytest.mark.parametrize("case", TRACED)
def test_c_oracle_vs_reference_traced_newton(case):
    def run(values, P, kw, xi_prev, grad_u):
        prob = oc.describe(values, P.active_idx, newton_mode="traced", strain_comps=9, **kw)
        return oc.mp_update(prob, xi_prev, grad_u, want=WANT)
    _check_traced(case, run)


@pytest.mark.parametrize("case", ["J2.mp", "hill_rot.fe", "hosford.mp"])
def test_torch_oracle_vs_reference_traced_newton(case):
    """The line-by-line torch-AD oracle on a few points of the same fixtures."""
    import torch
    from oracle import cmad_oracle as co
    kind, key, values, P = _setup(case)
    kw = dict(NEWTON[key])
    spec = co.ModelSpec()
    tv = co.to_torch_tree(values)
    for s in (0, TR[f"{case}.xi"].shape[0] - 1):
        for i in range(4):
            gu = torch.from_numpy(TR[f"{case}.grad_u"][s, i].reshape(3, 3).copy())
            x, info = co.newton_traced(TR[f"{case}.xi_prev"][s, i].copy(), tv, gu, gu, spec, **kw)
            assert info.iters == TR[f"{case}.iters"][s, i]
            assert (info.flag_entry | 2 * info.flag_exit) == TR[f"{case}.flags"][s, i]
            assert rel_err(x.numpy(), TR[f"{case}.xi"][s, i]) < 1e-10


@pytest.mark.parametrize("case", IMPER)
def test_c_oracle_vs_reference_imperative_newton(case):
    kind = case.split(".")[0]
    values = material(kind)
    prob = oc.describe(values, [], newton_mode="imperative", strain_comps=9)
    F = IM[f"{case}.F"]
    xi = np.zeros((7, 1))
    for t in range(1, F.shape[2]):
        gu = (F[:, :, t] - np.eye(3)).reshape(9, 1)
        r = oc.mp_update(prob, xi, gu, want=("xi", "sigma", "iters", "cnorm"))
        assert int(r["iters"][0]) == IM[f"{case}.iters"][t], (case, t)
        assert rel_err(r["xi"][:, 0], IM[f"{case}.xi"][t]) < 1e-10, (case, t)
        assert rel_err(r["sigma"][:, 0], IM[f"{case}.sigma"][t]) < 1e-10, (case, t)
        assert abs(r["cnorm"][0] - IM[f"{case}.cnorm"][t]) < 1e-11
        xi = r["xi"]


def _objective_inputs(case):
    kind, mode = case.split(".")
    values, act, tr = objective_trees(kind, mode == "scaled")
    P = Parameters(values, act, tr)
    assert np.array_equal(P.active_idx, OB[f"{case}.active_idx"])
    P.set_active_values_from_flat(OB[f"{case}.x_canonical"], are_canonical=True)
    assert np.allclose(P.flat_active_values(False), OB[f"{case}.active_native"], rtol=1e-14)
    F, data, w = OB[f"{case}.F"], OB[f"{case}.data"], OB[f"{case}.weight"]
    sh = np.ascontiguousarray((F - np.eye(3)[:, :, None]).reshape(9, 1, -1).transpose(2, 0, 1))   # (N+1, 9, 1)
    dh = np.ascontiguousarray(data.reshape(9, 1, -1).transpose(2, 0, 1))
    return P, F, data, w, sh, dh


@pytest.mark.parametrize("case", OBJ)
@pytest.mark.parametrize("strategy", ["adjoint", "direct"])
def test_oracle_objective_vs_reference(case, strategy):
    P, F, data, w, sh, dh = _objective_inputs(case)
    J, g, *_ = mo.objective(P.values, P.active_idx, sh, dh, w, strategy)
    P.transform_grad(g)
    assert abs(J - OB[f"{case}.J_{strategy}"]) < 1e-11 * abs(J)
    assert rel_err(g, OB[f"{case}.grad_{strategy}"]) < 1e-9


# ------------------------------------------------------------------------------------------ #
#  GPU: the CUDA path through the C-ABI against the reference's own output                   #
# ------------------------------------------------------------------------------------------ #
@pytest.mark.gpu
@pytest.mark.parametrize("case", TRACED)
@pytest.mark.parametrize("force_generic", [False, True])
def test_cuda_vs_reference_traced_newton(cuda_device, case, force_generic):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp

    def run(values, P, kw, xi_prev, grad_u):
        nw = NewtonSettings(mode="traced", force_generic=force_generic, **kw)
        out = mp.mp_update(material_from_values(values), nw, active_param_ids(P),
                           torch.from_numpy(xi_prev).to(cuda_device), torch.from_numpy(grad_u).to(cuda_device),
                           outputs=WANT[:-2])
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}
    if force_generic and not case.startswith("J2"):
        pytest.skip("only J2 has a specialised kernel to bypass")
    _check_traced(case, run)


@pytest.mark.gpu
@pytest.mark.parametrize("case", IMPER)
def test_cuda_vs_reference_imperative_newton(cuda_device, case):
    import torch
    from cmad_b200 import NewtonSettings, material_from_values, mp
    kind = case.split(".")[0]
    mat = material_from_values(material(kind))
    nw = NewtonSettings(mode="imperative")
    F = IM[f"{case}.F"]
    xi = torch.zeros((7, 1), dtype=torch.float64, device=cuda_device)
    for t in range(1, F.shape[2]):
        gu = torch.from_numpy((F[:, :, t] - np.eye(3)).reshape(9, 1)).to(cuda_device)
        out = mp.mp_update(mat, nw, np.zeros(0, np.int32), xi, gu, outputs=("xi", "sigma", "iters", "cnorm"))
        assert int(out["iters"][0]) == IM[f"{case}.iters"][t], (case, t)
        assert rel_err(out["xi"][:, 0].cpu().numpy(), IM[f"{case}.xi"][t]) < 1e-10, (case, t)
        assert rel_err(out["sigma"][:, 0].cpu().numpy(), IM[f"{case}.sigma"][t]) < 1e-10, (case, t)
        assert abs(float(out["cnorm"][0]) - IM[f"{case}.cnorm"][t]) < 1e-11
        xi = out["xi"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", OBJ)
def test_cuda_objectives_vs_reference(cuda_device, case):
    """The reference's constructor signatures, `evaluate(x_canonical)`, K1 + K2."""
    from cmad_b200.objectives import Calibration, MPAdjointObjective, MPDirectObjective, SmallElasticPlastic
    for strategy, ctor in (("adjoint", MPAdjointObjective), ("direct", MPDirectObjective)):
        P, F, data, w, _, _ = _objective_inputs(case)
        obj = ctor(Calibration(SmallElasticPlastic(P), data, w), F, device=cuda_device)
        r = obj.evaluate(OB[f"{case}.x_canonical"])
        assert abs(r.J - OB[f"{case}.J_{strategy}"]) < 1e-11 * abs(r.J)
        assert rel_err(r.grad, OB[f"{case}.grad_{strategy}"]) < 1e-9


def _leaf_parameters(kind):
    """Only the leaves the closed-form kernels did not differentiate in round 1 active: the 9
    rotation-matrix entries and (Hosford) the exponent a."""
    values = material(kind)
    act = const_like(values, False)
    act["rotation matrix"] = True
    if kind.startswith("hosford"):
        act["plastic"]["effective stress"]["hosford"]["a"] = True
    return values, Parameters(values, act, const_like(values, None))


def _check_leaf_columns(case, run_update):
    kind, key = case.split(".")
    values, P = _leaf_parameters(kind)
    aidx = np.asarray(P.active_idx)
    assert len(aidx) == 9 + (1 if kind.startswith("hosford") else 0)
    g = {k: TR[f"{case}.{k}"] for k in ("grad_u", "xi_prev", "iters", "flags", "dC_dp")}
    n_plastic = 0
    for s in range(g["grad_u"].shape[0]):
        out = run_update(values, P, _newton_kw(key), g["xi_prev"][s].T.copy(), g["grad_u"][s].T.copy())
        same = (np.asarray(out["iters"]) == g["iters"][s]) & (np.asarray(out["flags"]) == g["flags"][s])
        assert same.all(), (case, s)
        n = same.size
        ref = g["dC_dp"][s][:, :, aidx]                                  # (n, 7, P_a): the reference's jacrev columns
        got = np.asarray(out["dC_dp"]).T.reshape(n, 7, len(aidx))
        assert rel_err(got, ref) < 1e-9, (case, s, rel_err(got, ref))
        n_plastic += int((g["flags"][s] & 2).astype(bool).sum())
        assert np.abs(ref).max() > 0 or n_plastic == 0
    assert n_plastic > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["J2.mp", "hill_rot.fe", "hosford.mp"])
def test_cuda_dC_dp_rotation_and_exponent_leaves_vs_reference(cuda_device, case):
    """dC/dp columns of the `rotation matrix` entries and of `hosford a` (the reference's jacrev over
    ALL leaves, cmad/models/model.py:126-133) from K1, against the reference's own run."""
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp

    def run(values, P, kw, xi_prev, grad_u):
        out = mp.mp_update(material_from_values(values), NewtonSettings(mode="traced", **kw), active_param_ids(P),
                           torch.from_numpy(xi_prev).to(cuda_device), torch.from_numpy(grad_u).to(cuda_device),
                           outputs=("xi", "iters", "flags", "dC_dp"))
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}
    _check_leaf_columns(case, run)


@pytest.mark.parametrize("case", ["J2.mp", "hill_rot.fe", "hosford.mp"])
def test_c_oracle_dC_dp_rotation_and_exponent_leaves_vs_reference(case):
    def run(values, P, kw, xi_prev, grad_u):
        prob = oc.describe(values, P.active_idx, newton_mode="traced", strain_comps=9, **kw)
        return oc.mp_update(prob, xi_prev, grad_u, want=("xi", "iters", "flags", "dC_dp"))
    _check_leaf_columns(case, run)


def test_reference_counts_of_the_notch_material_are_stable_under_one_ulp():
    """tests/golden/make_knife_edge_golden.py: the reference's own make_newton_solve on the
    hosford_notch.notch inputs, re-run with every strain entry moved by +-1 ulp - same counts, same
    flags.  So "bit-exact counts" is a well-defined target for this material too."""
    K = np.load(os.path.join(G, "ref_knife_edge.npz"))
    assert K["iters"].shape == (2, 16) and int(K["iters"].max()) >= 10
    for tag in ("base", "up", "down"):
        assert np.array_equal(K[f"iters_{tag}"], K["iters"]) and np.array_equal(K[f"flags_{tag}"], K["flags"])
    assert np.array_equal(K["iters"], TR["hosford_notch.notch.iters"])
