"""Imperative `newton_solve` with its legacy line search (`max_ls_evals > 0`,
cmad/models/nonlinear_solver.py:55-81) against golden vectors produced by executing the reference's
own source (tests/golden/make_legacy_ls_golden.py): histories of the near-Tresca notch material
on which full Newton steps overshoot - without the search the reference runs into max_iters,
with it it converges in 8-10 updates.  Counts exact, state / stress 1e-9."""
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from tests.golden.materials import material
from tests.helpers import rel_err

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_imperative_legacy_ls.npz"))
CASES = [(n, ls) for n in ("notch_a", "notch_b", "J2") for ls in (0, 3, 8)]
KW = dict(max_iters=40, abs_tol=1e-14, rel_tol=1e-14)


def _walk(name, ls, update):
    F = G[f"{name}.F"]
    ref = {k: G[f"{name}.ls{ls}.{k}"] for k in ("xi", "sigma", "iters", "cnorm")}
    xi = np.zeros((7, 1))
    converged_steps = 0
    for t in range(1, F.shape[2]):
        gu = (F[:, :, t] - np.eye(3)).reshape(9, 1)
        out = update(xi, gu)
        assert int(out["iters"][0]) == ref["iters"][t], (name, ls, t, int(out["iters"][0]), ref["iters"][t])
        if ref["cnorm"][t] < 1e-12:           # converged steps: values; a step that ran into max_iters is chaotic
            assert rel_err(np.asarray(out["xi"])[:, 0], ref["xi"][t]) < 1e-9, (name, ls, t)
            assert rel_err(np.asarray(out["sigma"])[:, 0], ref["sigma"][t]) < 1e-9, (name, ls, t)
            converged_steps += 1
            xi = np.asarray(out["xi"]).copy()
        else:
            xi = ref["xi"][t].reshape(7, 1).copy()    # continue from the reference's state
    return converged_steps


def _converged_in_reference(name, ls):
    return int(np.sum(G[f"{name}.ls{ls}.cnorm"][1:] < 1e-12))


def test_fixture_shows_an_active_search():
    assert all(_converged_in_reference(n, 8) >= 4 for n in ("notch_a", "notch_b", "J2"))
    assert G["notch_a.ls0.iters"].max() == 40 and G["notch_a.ls8.iters"].max() <= 12
    assert G["notch_b.ls0.iters"].max() == 40 and G["notch_b.ls8.cnorm"].max() < 1e-14
    assert np.array_equal(G["J2.ls0.iters"], G["J2.ls8.iters"])


@pytest.mark.parametrize("name,ls", CASES)
def test_c_oracle_legacy_line_search_vs_reference(name, ls):
    values = material(str(G[f"{name}.kind"]))
    prob = oc.describe(values, [], newton_mode="imperative", strain_comps=9, max_ls_evals=ls, **KW)
    n = _walk(name, ls, lambda xi, gu: oc.mp_update(prob, xi, gu, want=("xi", "sigma", "iters", "cnorm")))
    assert n == _converged_in_reference(name, ls)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ls", CASES)
def test_cuda_legacy_line_search_vs_reference(cuda_device, name, ls):
    import torch
    from cmad_b200 import NewtonSettings, material_from_values, mp
    values = material(str(G[f"{name}.kind"]))
    mat = material_from_values(values)
    nw = NewtonSettings(mode="imperative", max_ls_evals=ls, **KW)

    def update(xi, gu):
        out = mp.mp_update(mat, nw, np.zeros(0, np.int32), torch.from_numpy(xi).to(cuda_device),
                           torch.from_numpy(gu).to(cuda_device), outputs=("xi", "sigma", "iters", "cnorm"))
        return {k: v.cpu().numpy() for k, v in out.items()}
    assert _walk(name, ls, update) == _converged_in_reference(name, ls)
