"""`cmad primal` on the B200 path (cmad_b200/primal.py): the forward loop's return tuple and
the on-disk layouts of cmad/io/writers.py:63-172 (cauchy (3,3,N+1), xi_block_<k> (N+1, n_eqs),
solver.json, J.json, grad / hess)."""
import json
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from oracle import analytic


def test_writers_produce_the_reference_layouts(tmp_path):
    from cmad_b200 import primal
    N = 5
    rng = np.random.default_rng(0)
    cauchy = rng.standard_normal((3, 3, N + 1))
    traj = [[rng.standard_normal(6), rng.standard_normal(1)] for _ in range(N + 1)]
    log = [{"iters": 3, "final_residual": 1e-16}] * N
    for fmt in ("npy", "text"):
        primal.write_cauchy(tmp_path, "run_", cauchy, fmt)
        primal.write_xi(tmp_path, "run_", traj, fmt)
        primal.write_grad(tmp_path, "run_", np.arange(3.0), fmt)
        primal.write_hessian(tmp_path, "run_", np.eye(3), fmt)
    primal.write_solver_log(tmp_path, "run_", log)
    primal.write_J(tmp_path, "run_", 1.25)
    assert np.array_equal(np.load(tmp_path / "run_cauchy.npy"), cauchy)
    b0, b1 = np.load(tmp_path / "run_xi_block_00.npy"), np.load(tmp_path / "run_xi_block_01.npy")
    assert b0.shape == (N + 1, 6) and b1.shape == (N + 1, 1)
    assert np.array_equal(b0[2], traj[2][0]) and np.array_equal(b1[4], traj[4][1])
    txt = np.loadtxt(tmp_path / "run_cauchy.csv")
    assert txt.shape == (N + 1, 9) and np.allclose(txt[3], cauchy[:, :, 3].reshape(9))
    assert open(tmp_path / "run_cauchy.csv").readline().strip() == "# S11 S12 S13 S21 S22 S23 S31 S32 S33"
    assert np.loadtxt(tmp_path / "run_xi_block_00.csv").shape == (N + 1, 6)
    assert json.load(open(tmp_path / "run_solver.json"))[0] == {"iters": 3, "final_residual": 1e-16}
    assert json.load(open(tmp_path / "run_J.json")) == {"J": 1.25}
    assert np.array_equal(np.load(tmp_path / "run_hess.npy"), np.eye(3))
    assert np.array_equal(np.loadtxt(tmp_path / "run_grad.csv"), np.arange(3.0))
    with pytest.raises(ValueError):
        primal.write_cauchy(tmp_path, "", cauchy, "hdf5")


@pytest.mark.gpu
def test_run_primal_pass_known_answer_and_oracle(cuda_device, tmp_path):
    """KA7 (tests/cli/test_primal_roundtrip.py:62-65): the uniaxial J2+Voce primal pass against
    the analytic path (rtol 1e-6), and against the torch oracle's run_primal_pass restatement
    (states 1e-10, Newton counts exact); then the files."""
    from cmad_b200 import primal
    from cmad_b200.objectives import Calibration, SmallElasticPlastic
    from oracle import cmad_oracle as co
    values, act, tr = analytic.j2_voce_param_tree("J2")
    mask = analytic.stress_masks_3d()[0]
    stress, strain, alpha = analytic.plastic_fields(mask, num_steps=30)
    F = analytic.deformation_gradient_history(strain)                 # step 0 = identity prepended
    stress = np.concatenate([np.zeros((3, 3, 1)), stress], axis=2)
    alpha = np.r_[0.0, alpha]
    N = F.shape[2] - 1
    model = SmallElasticPlastic(Parameters(values, act, tr))
    w = np.ones((3, 3))
    cauchy, traj, log, J = primal.run_primal_pass(model, F, N, {"max_iters": 10}, Calibration(model, stress, w),
                                                  device=cuda_device)
    assert cauchy.shape == (3, 3, N + 1) and len(traj) == N + 1 and len(log) == N
    assert np.abs(cauchy - stress).max() < 1e-6 * np.abs(stress).max()
    assert np.abs(np.array([t[1][0] for t in traj]) - alpha).max() < 1e-6
    assert J < 1e-10 * 0.5 * (stress ** 2).sum()
    xi_o, cauchy_o, iters_o, norms_o, _ = co.mp_primal(co.OracleParameters(values, act, tr), F, co.ModelSpec())
    assert np.abs(cauchy - cauchy_o).max() < 1e-10 * np.abs(cauchy_o).max()
    got = np.array([np.concatenate(t) for t in traj])
    assert np.abs(got - xi_o).max() < 1e-10 * np.abs(xi_o).max()
    assert [s["iters"] for s in log] == list(iters_o[1:])
    primal.write_cauchy(tmp_path, "", cauchy, "npy"); primal.write_xi(tmp_path, "", traj, "npy")
    primal.write_solver_log(tmp_path, "", log)
    assert np.load(tmp_path / "cauchy.npy").shape == (3, 3, N + 1)
    assert np.load(tmp_path / "xi_block_00.npy").shape == (N + 1, 6)
    assert np.load(tmp_path / "xi_block_01.npy").shape == (N + 1, 1)
    assert len(json.load(open(tmp_path / "solver.json"))) == N
    # a batch of points = the same pass with a leading axis
    Fb = np.stack([F, np.eye(3)[:, :, None] + 0.5 * (F - np.eye(3)[:, :, None])])
    cb, tb, lb, _ = primal.run_primal_pass(model, Fb, N, None, device=cuda_device)
    assert cb.shape == (2, 3, 3, N + 1) and np.array_equal(cb[0], cauchy)
    assert tb[5][0].shape == (2, 6) and lb[0]["iters"].shape == (2,)


@pytest.mark.gpu
@pytest.mark.parametrize("dtn", ["PLANE_STRESS", "UNIAXIAL_STRESS"])
def test_run_primal_pass_def_types_vs_reference(cuda_device, dtn):
    """The def-type variants against the reference's own imperative run (ref_def_types.npz):
    three residual blocks, the stretch block last."""
    from cmad_b200 import objectives as ob, primal
    from tests.golden.materials import const_like, material
    DT = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_def_types.npz"))
    case = f"J2.{dtn}"
    values = material("J2")
    model = ob.SmallElasticPlastic(Parameters(values, const_like(values, False), const_like(values, None)),
                                   def_type=getattr(ob, dtn))
    F = DT[f"{case}.F"]
    N = F.shape[2] - 1
    cauchy, traj, log, _ = primal.run_primal_pass(model, F, N, None, device=cuda_device)
    assert len(traj[0]) == 3 and traj[0][2].shape == ((1,) if dtn == "PLANE_STRESS" else (2,))
    assert np.all(traj[0][2] == 1.0)
    got = np.array([np.concatenate(t) for t in traj[1:]])
    assert np.abs(got - DT[f"{case}.xi"]).max() < 1e-10 * np.abs(DT[f"{case}.xi"]).max()
    assert [s["iters"] for s in log] == list(DT[f"{case}.iters"])
    from tests.helpers import UP
    sig6 = np.array([[cauchy[i, j, t] for i, j in UP] for t in range(1, N + 1)])
    assert np.abs(sig6 - DT[f"{case}.sigma"]).max() < 1e-10 * np.abs(DT[f"{case}.sigma"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["J2", "hill"])
def test_run_primal_pass_rate_model_vs_reference(cuda_device, kind):
    """`cmad primal` with SmallRateElasticPlastic (state = [cauchy, alpha]) against the
    reference's own imperative run (ref_rate_model.npz): states, stresses, Newton counts."""
    from cmad_b200 import objectives as ob, primal
    from tests.golden.materials import const_like, material
    RT = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rate_model.npz"))
    if f"{kind}.F" not in RT.files:
        pytest.skip("no fixture for this surface")
    values = material(kind)
    model = ob.SmallRateElasticPlastic(Parameters(values, const_like(values, False), const_like(values, None)))
    F = RT[f"{kind}.F"]
    N = F.shape[2] - 1
    cauchy, traj, log, _ = primal.run_primal_pass(model, F, N, None, device=cuda_device)
    assert len(traj[0]) == 2 and traj[0][0].shape == (6,)
    got = np.array([np.concatenate(t) for t in traj[1:]])
    assert np.abs(got - RT[f"{kind}.xi"]).max() < 1e-10 * np.abs(RT[f"{kind}.xi"]).max()
    assert [s["iters"] for s in log] == list(RT[f"{kind}.iters"])
    from tests.helpers import UP
    sig6 = np.array([[cauchy[i, j, t] for i, j in UP] for t in range(1, N + 1)])
    assert np.abs(sig6 - RT[f"{kind}.sigma"]).max() < 1e-10 * np.abs(RT[f"{kind}.sigma"]).max()
