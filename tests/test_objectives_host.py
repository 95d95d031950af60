"""CPU tests of the objective host logic: the batched NumPy/C++ oracle objective
against the line-by-line torch oracle, the reference's own gradient checks
(adjoint == direct, finite-difference error drop; tests/objectives/
test_J2_fd_checks.py:303-386), and the world_size-2 gloo path of
BatchedMPObjective (sharding + the single allreduce + transform_grad)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmad_b200 import Parameters
from cmad_b200.objectives import BatchedMPObjective, shard_range
from oracle import analytic, cmad_oracle as co, mp_objective_np as mo
from tests.helpers import param_tree, random_strains


def _problem(n=6, N=10, seed=0, kind="J2"):
    rng = np.random.default_rng(seed)
    d = random_strains(rng, n, scale=1.0, diag_only=(kind == "hosford"))
    d /= np.sqrt((d ** 2).sum(0))
    amp = rng.uniform(2e-3, 5e-3, size=n)
    sh = np.zeros((N + 1, 6, n))
    for t in range(1, N + 1):
        sh[t] = d * amp * (t / N)
    data = rng.normal(size=(N + 1, 9, n)) * 50
    w = np.array([[1, 0.5, 0], [0.5, 1, 0], [0, 0, 0.3]])
    return sh, data, w


def test_numpy_objective_matches_torch_oracle_per_point():
    values, act, tr = param_tree("J2")
    P = co.OracleParameters(values, act, tr)
    sh, data, w = _problem(n=2, N=8)
    J, g, Jp, gp, xi, it = mo.objective(values, P.active_idx, sh, data, w, "adjoint")
    _, _, _, gd, _, _ = mo.objective(values, P.active_idx, sh, data, w, "direct")
    assert np.allclose(gp, gd, rtol=1e-10, atol=1e-10 * np.abs(gp).max())     # adjoint == direct
    for i in range(2):
        F = np.repeat(np.eye(3)[:, :, None], sh.shape[0], axis=2)
        for t in range(sh.shape[0]):
            e = sh[t, :, i]
            F[:, :, t] += np.array([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]])
        dat = data[:, :, i].T.reshape(3, 3, -1)
        Jt, gt = co.mp_objective_adjoint(co.OracleParameters(values, act, tr), F, dat, w, co.ModelSpec())
        assert abs(Jt - Jp[i]) < 1e-12 * abs(Jt)
        assert np.abs(gt - gp[:, i]).max() < 1e-9 * np.abs(gt).max()


def test_gradient_finite_difference_error_drop():
    """Directional FD check as the reference does it: the error must fall by > 5
    decades as h shrinks (test_J2_fd_checks.py:338-349)."""
    values, act, tr = param_tree("J2", active=("D", "S", "Y"))
    P = co.OracleParameters(values, act, tr)
    sh, data, w = _problem(n=3, N=10, seed=3)
    x0 = 1.1 * P.flat_active_values()

    def J_of(x):
        Pn = co.OracleParameters(param_tree("J2", active=("D", "S", "Y"))[0], act, tr)
        Pn.set_active_values_from_flat(x, are_canonical=False)
        return mo.objective(Pn.values, Pn.active_idx, sh, data, w, "adjoint")[:2]

    J0, g0 = J_of(x0)
    direction = np.array([0.3, -0.5, 0.8]) * x0
    errs = []
    for h in np.logspace(-1, -7, 7):
        fd = (J_of(x0 + h * direction)[0] - J_of(x0 - h * direction)[0]) / (2 * h)
        errs.append(abs(fd - g0 @ direction))
    assert np.log10(max(errs) / min(errs)) > 5.0


def _oracle_local_evaluator(parameters, values_template, sh, data, w, strategy):
    def ev():
        J, g, *_ = mo.objective(parameters.values, parameters.active_idx, sh, data, w, strategy)
        return torch.tensor(np.concatenate([[J], g]))
    return ev


def _gloo_worker(rank, world, port, sh, data, w, x, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    values, act, tr = analytic.j2_voce_param_tree("J2")      # log/bounds transforms active
    P = Parameters(values, act, tr)
    lo, hi = shard_range(sh.shape[2], rank, world)
    ev = _oracle_local_evaluator(P, values, sh[:, :, lo:hi], data[:, :, lo:hi], w, "adjoint")
    from cmad_b200.comm import WORLD
    res = BatchedMPObjective(P, ev, group=WORLD).evaluate(x)
    # an UNSHARDED objective (every rank evaluates the same experiment, as the reference-signature
    # constructors do) must not be summed over the ranks of an initialised job
    ev_all = _oracle_local_evaluator(P, values, sh, data, w, "adjoint")
    res_all = BatchedMPObjective(P, ev_all).evaluate(x)
    if rank == 0:
        ret["J"], ret["grad"] = res.J, res.grad
        ret["J_unsharded"], ret["grad_unsharded"] = res_all.J, res_all.grad
    dist.destroy_process_group()


def test_two_rank_gloo_objective_equals_single_process():
    sh, data, w = _problem(n=7, N=6, seed=5)                 # 7 points: ragged 4 + 3 split
    x = np.array([0.1, -0.2, 0.05])                          # canonical coordinates
    values, act, tr = analytic.j2_voce_param_tree("J2")
    P = Parameters(values, act, tr)
    single = BatchedMPObjective(P, _oracle_local_evaluator(P, values, sh, data, w, "adjoint")).evaluate(x)
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, sh, data, w, x, ret), nprocs=2, join=True)
    assert abs(ret["J"] - single.J) < 1e-12 * abs(single.J)
    assert np.allclose(ret["grad"], single.grad, rtol=1e-11, atol=0)
    assert ret["J_unsharded"] == single.J and np.array_equal(ret["grad_unsharded"], single.grad)
    # the canonical chain rule was applied once, after the reduction
    Pn = Parameters(*analytic.j2_voce_param_tree("J2")); Pn.set_active_values_from_flat(x)
    _, g_native, *_ = mo.objective(Pn.values, Pn.active_idx, sh, data, w, "adjoint")
    g_c = g_native.copy(); Pn.transform_grad(g_c)
    assert np.allclose(g_c, single.grad, rtol=1e-12)


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
