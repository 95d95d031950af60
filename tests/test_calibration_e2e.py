"""End-to-end calibration on the B200 path (what `cmad calibrate` does with scipy on top of
`MPObjective.evaluate`, cmad/cli/calibrate.py): synthetic stress data generated at the "true"
parameters by the primal pass, then (i) L-BFGS-B on J with the adjoint gradient and (ii) plain
Newton iterations with the K2-H Hessian, both in canonical coordinates (log / bounds transforms),
recover the parameters.  The path under test: run_primal_pass -> fused forward history -> K2
adjoint -> K2-H -> Parameters.transform_grad / transform_hessian."""
import numpy as np
import pytest

from cmad_b200 import Parameters
from oracle import analytic

pytestmark = pytest.mark.gpu


def _setup(cuda_device, n_points=6, N=20):
    from cmad_b200 import primal, synthetic
    from cmad_b200.objectives import (BatchedMPObjective, SmallElasticPlastic, data_history, gpu_local_evaluator,
                                      strain_history_from_F)
    values, act, tr = analytic.j2_voce_param_tree("J2")              # Y (log), S, D (bounds) active
    P = Parameters(values, act, tr)
    model = SmallElasticPlastic(P)
    d, d2, a = synthetic.path_params(5, 0, n_points)
    F = np.repeat(np.eye(3)[None, :, :, None], n_points, axis=0).repeat(N + 1, axis=3)
    for t in range(1, N + 1):
        e = synthetic.strain_at_step(d, d2, 1.5 * a, int(round(t * 100 / N)))       # (6, n)
        for p in range(n_points):
            F[p, :, :, t] += np.array([[e[0, p], e[1, p], e[2, p]], [e[1, p], e[3, p], e[4, p]],
                                       [e[2, p], e[4, p], e[5, p]]])
    x_true = P.flat_active_values(True).copy()
    native_true = P.flat_active_values(False).copy()
    cauchy, _, log, _ = primal.run_primal_pass(model, F, N, None, device=cuda_device)
    assert max(int(s["iters"].max()) for s in log) >= 2                       # plastic loading
    w = np.array([[1.0, 0.5, 0.5], [0.5, 1.0, 0.5], [0.5, 0.5, 1.0]])
    sh, dh = strain_history_from_F(F), data_history(cauchy)
    grad_obj = BatchedMPObjective(P, gpu_local_evaluator(model, sh, dh, w, "adjoint", cuda_device))
    hess_obj = BatchedMPObjective(P, gpu_local_evaluator(model, sh, dh, w, "direct_adjoint", cuda_device))
    return P, grad_obj, hess_obj, x_true, native_true


def test_lbfgs_with_the_adjoint_gradient_recovers_the_parameters(cuda_device):
    from scipy.optimize import minimize
    P, grad_obj, _, x_true, native_true = _setup(cuda_device)
    assert grad_obj.evaluate(x_true).J < 1e-18 * 1e6                          # data reproduce at the truth
    P.set_active_values_from_flat(1.15 * native_true, False)
    x0 = P.flat_active_values(True).copy()
    J0 = grad_obj.evaluate(x0).J

    def fun(x):
        r = grad_obj.evaluate(x)
        return r.J, r.grad
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", options={"maxiter": 200, "ftol": 1e-20, "gtol": 1e-12})
    assert res.fun < 1e-10 * J0
    P.set_active_values_from_flat(res.x, True)
    assert np.allclose(P.flat_active_values(False), native_true, rtol=1e-5)


def test_newton_with_the_k2h_hessian_converges_quadratically(cuda_device):
    P, grad_obj, hess_obj, x_true, native_true = _setup(cuda_device)
    P.set_active_values_from_flat(1.1 * native_true, False)
    x = P.flat_active_values(True).copy()
    gnorms = []
    for _ in range(8):
        r = hess_obj.evaluate(x)
        gnorms.append(float(np.linalg.norm(r.grad)))
        x = x - np.linalg.solve(r.hessian, r.grad)
    # Newton, not gradient descent: an order of magnitude or more per step from the start, until
    # the elastic / plastic branch switches of individual steps (J is only piecewise smooth in
    # the yield parameters) take over near the optimum
    assert gnorms[1] < 0.2 * gnorms[0] and gnorms[2] < 0.05 * gnorms[1] and gnorms[3] < 0.05 * gnorms[2], gnorms
    assert gnorms[-1] < 1e-6 * gnorms[0], gnorms
    P.set_active_values_from_flat(x, True)
    assert np.allclose(P.flat_active_values(False), native_true, rtol=1e-6)
