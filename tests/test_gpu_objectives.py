"""GPU parity of K2 (adjoint / direct calibration gradients over whole load
histories) against the CPU oracle: J to 1e-12, gradients to 1e-9 relative (the
reference's own cross-strategy tolerances, tests/objectives/
test_jvp_vs_original.py:95-97), adjoint == direct, run-to-run bit reproducibility."""
import numpy as np
import pytest
import torch

from cmad_b200 import Parameters
from cmad_b200.objectives import (BatchedMPObjective, Calibration, MPAdjointObjective,
                                  MPDirectObjective, SmallElasticPlastic, gpu_local_evaluator)
from oracle import analytic, mp_objective_np as mo
from tests.helpers import param_tree
from tests.test_objectives_host import _problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
def test_batched_objective_matches_oracle(cuda_device, kind):
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    # Hosford: the exponent a is an active parameter too (d/da of |x|^a and of the outer 1/a root)
    active = ("E", "nu", "D", "S", "Y") + (tuple("FGHLMN") if kind == "hill" else ()) + (("a",) if kind == "hosford" else ())
    values, act, tr = param_tree(kind, ("voce",), hill=hill, active=active)
    P = Parameters(values, act, tr)
    assert len(P.active_idx) == len(active)
    sh, data, w = _problem(n=3000, N=12, seed=1, kind=kind)
    Jr, gr, Jp, gp, xi_ref, it_ref = mo.objective(values, P.active_idx, sh, data, w, "adjoint")
    model = SmallElasticPlastic(P)
    res = {}
    for strategy in ("adjoint", "direct"):
        ev = gpu_local_evaluator(model, sh, data, w, strategy, cuda_device)
        out = ev().cpu().numpy()
        res[strategy] = out
        assert abs(out[0] - Jr) < 1e-12 * abs(Jr)
        assert np.abs(out[1:] - gr).max() < 1e-9 * np.abs(gr).max()
        h = ev.histories
        assert np.array_equal(h.iters.cpu().numpy(), it_ref)              # Newton counts, every step
        assert np.abs(h.xi.cpu().numpy() - xi_ref).max() < 1e-10 * np.abs(xi_ref).max()
        assert np.abs(h.J_point.cpu().numpy() - Jp).max() < 1e-12 * np.abs(Jp).max()
        again = ev().cpu().numpy()
        assert np.array_equal(out, again)                                  # deterministic reduction
    assert np.abs(res["adjoint"][1:] - res["direct"][1:]).max() < 1e-10 * np.abs(gr).max()


def test_reference_style_single_point_objectives(cuda_device):
    """The reference's constructor signatures on one material point with
    canonical (log / bounds) parameter transforms: adjoint == direct == oracle."""
    values, act, tr = analytic.j2_voce_param_tree("J2")
    sh, data, w = _problem(n=1, N=20, seed=2)
    F = np.repeat(np.eye(3)[:, :, None], 21, axis=2)
    for t in range(21):
        e = sh[t, :, 0]
        F[:, :, t] += np.array([[e[0], e[1], e[2]], [e[1], e[3], e[4]], [e[2], e[4], e[5]]])
    dat = data[:, :, 0].T.reshape(3, 3, 21)
    x = np.array([0.1, 0.1, 0.1])
    results = []
    for ctor in (MPAdjointObjective, MPDirectObjective):
        P = Parameters(*analytic.j2_voce_param_tree("J2"))
        obj = ctor(Calibration(SmallElasticPlastic(P), dat, w), F, device=cuda_device)
        results.append(obj.evaluate(x))
    Pn = Parameters(*analytic.j2_voce_param_tree("J2")); Pn.set_active_values_from_flat(x)
    Jr, gr, *_ = mo.objective(Pn.values, Pn.active_idx, sh, data, w, "adjoint")
    Pn.transform_grad(gr)
    for r in results:
        assert abs(r.J - Jr) < 1e-12 * abs(Jr)
        assert np.allclose(r.grad, gr, rtol=1e-9)
    assert np.allclose(results[0].grad, results[1].grad, rtol=1e-10)


@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
def test_fused_forward_history_equals_per_step_path(cuda_device, kind, monkeypatch):
    """cmadx_mp_forward_history: the fused one-launch kernel (state in registers across the load
    steps) against the per-step K1 launches it replaces: identical states and Newton counts."""
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    values, act, tr = param_tree(kind, ("voce",), hill=hill)
    P = Parameters(values, act, tr)
    sh, data, w = _problem(n=20000, N=15, seed=4, kind=kind)
    model = SmallElasticPlastic(P)
    res = {}
    for mode in ("fused", "per_step"):
        if mode == "per_step":
            monkeypatch.setenv("CMADX_HISTORY_PER_STEP", "1")
        ev = gpu_local_evaluator(model, sh, data, w, "adjoint", cuda_device)
        out = ev().cpu().numpy()
        h = ev.histories
        res[mode] = (out, h.xi.cpu().numpy().copy(), h.iters.cpu().numpy().copy())
    assert np.array_equal(res["fused"][2], res["per_step"][2])
    assert np.array_equal(res["fused"][1], res["per_step"][1])
    assert np.array_equal(res["fused"][0], res["per_step"][0])
    assert res["fused"][2].max() >= 2


@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
def test_batched_objective_rotated_material_axes(cuda_device, kind):
    """K2 with a rotated "rotation matrix" (cmad/models/small_elastic_plastic.py:44-62, 318-319):
    the state lives in material axes, the QoI compares the global cauchy Q sigma_m Q^T - adjoint
    and direct gradients against the oracle, for every yield surface."""
    from tests.helpers import rotation_matrix
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    active = ("E", "nu", "D", "S", "Y") + (tuple("FGHLMN") if kind == "hill" else ())
    values, act, tr = param_tree(kind, ("voce",), hill=hill, active=active,
                                 rotation=rotation_matrix([1.0, 2.0, -0.5], 0.7))
    P = Parameters(values, act, tr)
    sh, data, w = _problem(n=2000, N=10, seed=7, kind="J2")      # full (non-diagonal) strain paths
    Jr, gr, Jp, gp, xi_ref, it_ref = mo.objective(values, P.active_idx, sh, data, w, "adjoint")
    model = SmallElasticPlastic(P)
    res = {}
    for strategy in ("adjoint", "direct"):
        ev = gpu_local_evaluator(model, sh, data, w, strategy, cuda_device)
        out = ev().cpu().numpy()
        res[strategy] = out
        assert abs(out[0] - Jr) < 1e-12 * abs(Jr)
        assert np.abs(out[1:] - gr).max() < 1e-9 * np.abs(gr).max(), (strategy, out[1:], gr)
        h = ev.histories
        assert np.array_equal(h.iters.cpu().numpy(), it_ref)
        assert np.abs(h.xi.cpu().numpy() - xi_ref).max() < 1e-10 * np.abs(xi_ref).max()
    assert np.abs(res["adjoint"][1:] - res["direct"][1:]).max() < 1e-10 * np.abs(gr).max()
    # the rotation matters: same histories with identity axes give a different objective (Hill)
    if kind == "hill":
        v0, a0, t0 = param_tree(kind, ("voce",), hill=hill, active=active)
        J0 = mo.objective(v0, Parameters(v0, a0, t0).active_idx, sh, data, w, "adjoint")[0]
        assert abs(J0 - Jr) > 1e-6 * abs(Jr)


@pytest.mark.parametrize("kind", ["J2", "hill"])
def test_host_buffer_objective_matches_device_histories(cuda_device, kind):
    """cmadx_mp_objective_host (histories in host memory, chunked pipeline, 8 (1 + P_a) bytes back
    per chunk) vs the device-resident evaluator: J and the gradient agree to summation order,
    the per-point objective exactly; chunked and unchunked runs agree; deterministic."""
    from cmad_b200 import NewtonSettings, active_param_ids, mp
    hill = (0.45, 0.6, 0.55, 1.4, 1.6, 1.5) if kind == "hill" else None
    values, act, tr = param_tree(kind, ("voce",), hill=hill, active=("E", "nu", "D", "S", "Y"))
    P = Parameters(values, act, tr)
    sh, data, w = _problem(n=5003, N=10, seed=4, kind=kind)
    model = SmallElasticPlastic(P)
    nw = NewtonSettings(mode="imperative", max_iters=10, abs_tol=1e-14, rel_tol=1e-14)
    pid = active_param_ids(P)
    for strategy in ("adjoint", "direct"):
        ev = gpu_local_evaluator(model, sh, data, w, strategy, cuda_device)
        dev = ev().cpu().numpy()
        Jp_dev = ev.histories.J_point.cpu().numpy()
        one, Jp1 = mp.mp_objective_host(model.material(), nw, pid, sh, data, w, strategy, device=cuda_device.index or 0,
                                        chunk_points=1 << 20, want_J_point=True)
        many, Jp2 = mp.mp_objective_host(model.material(), nw, pid, sh, data, w, strategy, device=cuda_device.index or 0,
                                         chunk_points=1024, want_J_point=True)
        assert np.array_equal(Jp1, Jp_dev) and np.array_equal(Jp2, Jp_dev)
        for got in (one, many):
            assert abs(got[0] - dev[0]) < 1e-13 * abs(dev[0])
            assert np.abs(got[1:] - dev[1:]).max() < 1e-12 * np.abs(dev[1:]).max()
        again = mp.mp_objective_host(model.material(), nw, pid, sh, data, w, strategy, device=cuda_device.index or 0,
                                     chunk_points=1024)
        assert np.array_equal(again, many)
    empty = mp.mp_objective_host(model.material(), nw, pid, sh[:, :, :0].copy(), data[:, :, :0].copy(), w)
    assert np.array_equal(empty, np.zeros(1 + len(pid)))
