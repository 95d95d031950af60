"""Pin the CPU oracle on the reference's own known answers.

KA1: analytic J2+Voce proportional path (cmad/verification/solutions.py:30-58;
tests/models/test_elastic_plastic_models.py:15-125) for J2, Hill(0.5 x 6) and
Hosford(a=4), uniaxial and biaxial masks, 100 steps, tolerance 1e-6 on alpha,
stress and the calibration objective - the reference test's own assertions.
"""
import numpy as np
import pytest
import torch

from oracle import analytic, cmad_oracle as co, oracle_c as oc
from tests.helpers import UP


def _c_oracle_path(kind, mask, mode):
    values, active, transforms = analytic.j2_voce_param_tree(kind)
    stress, strain, alpha = analytic.plastic_fields(mask)
    prob = oc.describe(values, [], newton_mode=mode)
    xi = np.zeros((7, 1))
    alphas, sig, its = [], [], []
    for k in range(strain.shape[2]):
        e = np.array([[strain[i, j, k]] for i, j in UP])
        r = oc.mp_update(prob, xi, e, want=("xi", "sigma", "iters"))
        xi = r["xi"]
        alphas.append(xi[6, 0]); sig.append(r["sigma"][:, 0]); its.append(r["iters"][0])
    sig = np.array(sig)
    ref = np.array([[stress[i, j, k] for i, j in UP] for k in range(stress.shape[2])])
    return np.array(alphas), sig, ref, alpha, np.array(its)


@pytest.mark.parametrize("kind", ["J2", "hill", "hosford"])
@pytest.mark.parametrize("mask_id", [0, 1])
@pytest.mark.parametrize("mode", ["imperative", "traced"])
def test_ka1_analytic_path_c_oracle(kind, mask_id, mode):
    mask = analytic.stress_masks_3d()[mask_id]
    a, sig, ref, alpha, its = _c_oracle_path(kind, mask, mode)
    assert np.linalg.norm(a - alpha) < 1e-6
    w = np.array([1, 2, 2, 1, 2, 1])[None, :]          # Frobenius norm over the 3x3 tensor
    assert np.sqrt((w * (sig - ref) ** 2).sum()) < 1e-6
    # calibration objective with zero data and weight |mask| (reference :98-125)
    wt = np.array([abs(mask[i, j]) for i, j in UP])[None, :]
    J = 0.5 * (w * (wt * sig) ** 2).sum()
    assert abs(J - 0.5 * (w * (wt * ref) ** 2).sum()) < 1e-6 * max(1.0, J) or abs(J - 0.5 * (w * (wt * ref) ** 2).sum()) < 1e-6
    assert its.max() <= 10


def test_ka1_analytic_path_torch_oracle_j2_uniaxial():
    """Same check through the torch.func (AD) oracle on the imperative MP driver
    (cli/primal.py:129-176), shortened to 25 steps to keep the CPU suite fast."""
    values, active, transforms = analytic.j2_voce_param_tree("J2")
    P = co.OracleParameters(values, active, transforms)
    mask = analytic.stress_masks_3d()[0]
    stress, strain, alpha = analytic.plastic_fields(mask, num_steps=25, max_alpha=0.12)
    F = analytic.deformation_gradient_history(strain)
    xi, cauchy, iters, norms, flags = co.mp_primal(P, F, co.ModelSpec())
    assert np.linalg.norm(xi[1:, 6] - alpha) < 1e-6
    assert np.linalg.norm(cauchy[:, :, 1:] - stress) < 1e-6


def test_ka2_tet_fixture_tangent_vs_fd():
    """KA2 (tests/global_residuals/test_for_model_coupled.py:34-82, 231-295): tet
    barycentre shapes, U[1,0]=.005, U[2,1]=.003, U[3,2]=.002, xi_prev=0: plastic,
    ||C(xi*)|| < 1e-10, alpha > 0, IFT dR/dU vs central FD (eps 1e-6, rtol 1e-5,
    atol 1e-7)."""
    values, _, _ = analytic.j2_voce_param_tree("J2")
    params = co.to_torch_tree(values)
    spec = co.ModelSpec()
    grad_N = np.array([[-1., -1., -1.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    U = np.zeros((4, 3)); U[1, 0] = .005; U[2, 1] = .003; U[3, 2] = .002
    Up = np.zeros((4, 3))
    R, dR, x, info = co.coupled_ip(params, U, Up, np.zeros(7), grad_N, 1.0, 1.0 / 6.0, spec)
    gu = co.interpolate_grad_u(torch.as_tensor(U), torch.as_tensor(grad_N))
    C = co.sep_residual(x, torch.zeros(7, dtype=co.DT), params, gu, gu * 0, spec)
    assert float(torch.linalg.norm(C)) < 1e-10
    assert float(x[6]) > 0.0 and info.flag_exit == 1
    fd = np.zeros((4, 3, 4, 3)); eps = 1e-6
    for b in range(4):
        for k in range(3):
            Um = U.copy(); Um[b, k] -= eps; Upl = U.copy(); Upl[b, k] += eps
            Rp = co.coupled_ip(params, Upl, Up, np.zeros(7), grad_N, 1.0, 1.0 / 6.0, spec, want_tangent=False)[0]
            Rm = co.coupled_ip(params, Um, Up, np.zeros(7), grad_N, 1.0, 1.0 / 6.0, spec, want_tangent=False)[0]
            fd[:, :, b, k] = ((Rp - Rm) / (2 * eps)).numpy()
    assert np.allclose(dR.numpy(), fd, rtol=1e-5, atol=1e-7)
    assert R.shape == (4, 3) and dR.shape == (4, 3, 4, 3)


def test_ka4_elastic_coupled_equals_closed_form():
    """KA4 (test_for_model_coupled.py:193-220): Elastic(kappa=100, mu=50); the
    COUPLED local Newton reproduces the closed-form stress and tangent to 1e-12."""
    values = {"elastic": {"kappa": 100.0, "mu": 50.0}}
    params = co.to_torch_tree(values)
    spec = co.ModelSpec(kind="elastic")
    grad_N = np.array([[-1., -1., -1.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    U = np.zeros((4, 3)); U[1, 0] = 0.001; U[2, 1] = 0.0005
    R, dR, x, info = co.coupled_ip(params, U, U * 0, np.zeros(6), grad_N, 1.0, 1.0 / 6.0, spec)
    gN = torch.as_tensor(grad_N)
    def closed(Uf):
        gu = co.interpolate_grad_u(Uf.reshape(4, 3), gN)
        s = co.isotropic_linear_elastic_cauchy_stress(torch.eye(3, dtype=co.DT) + gu, params)
        return (gN @ s) / 6.0
    Rc = closed(torch.as_tensor(U).reshape(-1))
    dRc = torch.func.jacfwd(closed)(torch.as_tensor(U).reshape(-1)).reshape(4, 3, 4, 3)
    assert torch.allclose(R, Rc, rtol=0, atol=1e-12)
    assert torch.allclose(dR, dRc, rtol=0, atol=1e-12)
    assert info.iters == 1
