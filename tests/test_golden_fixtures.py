"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE
(tests/golden/make_golden.py runs cmad/verification/solutions.py:30-58 and
functions.py:7-54 from /root/reference, unmodified): the analytic J2+Voce
proportional paths of tests/models/test_elastic_plastic_models.py:15-125.

CPU (not gpu): the oracle's restatement of the generator is bit-close to the
reference's output, and the C++ oracle driven along the golden strain history
reproduces the golden stress / alpha at the reference test's own 1e-6.
GPU: the CUDA path (K1 through the C-ABI) does the same.
"""
import os

import numpy as np
import pytest

from oracle import analytic, oracle_c as oc
from tests.helpers import UP

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ka1_analytic_paths.npz"))
CASES = [(y, m) for y in ("J2", "hill") for m in ("uniaxial", "biaxial")]


def _vec6(T):          # (3,3,N) -> (N,6) in CMAD's packing
    return np.array([[T[i, j, k] for i, j in UP] for k in range(T.shape[2])])


@pytest.mark.parametrize("yname,mname", CASES)
def test_restated_generator_equals_reference_output(yname, mname):
    mask = GOLD[f"{yname}_{mname}_mask"]
    stress, strain, alpha = analytic.plastic_fields(mask)        # J2 restatement; Hill(0.5) == J2
    assert np.array_equal(alpha, GOLD[f"{yname}_{mname}_alpha"])
    assert np.abs(stress - GOLD[f"{yname}_{mname}_stress"]).max() < 1e-10
    assert np.abs(strain - GOLD[f"{yname}_{mname}_strain"]).max() < 1e-15


@pytest.mark.parametrize("yname,mname", CASES)
@pytest.mark.parametrize("mode", ["imperative", "traced"])
def test_c_oracle_reproduces_golden_path(yname, mname, mode):
    values, _, _ = analytic.j2_voce_param_tree(yname)
    prob = oc.describe(values, [], newton_mode=mode)
    strain, stress, alpha = (GOLD[f"{yname}_{mname}_{k}"] for k in ("strain", "stress", "alpha"))
    e6 = _vec6(strain)
    xi = np.zeros((7, 1)); a, s = [], []
    for k in range(e6.shape[0]):
        r = oc.mp_update(prob, xi, e6[k][:, None], want=("xi", "sigma"))
        xi = r["xi"]; a.append(xi[6, 0]); s.append(r["sigma"][:, 0])
    w = np.array([1, 2, 2, 1, 2, 1])[None, :]
    assert np.linalg.norm(np.array(a) - alpha) < 1e-6
    assert np.sqrt((w * (np.array(s) - _vec6(stress)) ** 2).sum()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("yname,mname", CASES + [("hosford", "uniaxial"), ("hosford", "biaxial")])
@pytest.mark.parametrize("mode", ["imperative", "traced"])
def test_cuda_path_reproduces_golden_path(cuda_device, yname, mname, mode):
    """K1 through the C-ABI along the reference's golden strain history: alpha and
    stress within the reference test's 1e-6 norms.  Hosford(a=4) shares the J2
    golden path, as in the reference's test (test_elastic_plastic_models.py:128-145)."""
    import torch
    from cmad_b200 import NewtonSettings, material_from_values, mp
    values, _, _ = analytic.j2_voce_param_tree(yname)
    gname = "J2" if yname == "hosford" else yname
    strain, stress, alpha = (GOLD[f"{gname}_{mname}_{k}"] for k in ("strain", "stress", "alpha"))
    mat = material_from_values(values)
    nw = NewtonSettings(mode=mode)
    e6 = torch.from_numpy(_vec6(strain)).to(cuda_device)
    xi = torch.zeros((7, 1), dtype=torch.float64, device=cuda_device)
    a, s = [], []
    for k in range(e6.shape[0]):
        out = mp.mp_update(mat, nw, np.zeros(0, np.int32), xi, e6[k][:, None].contiguous(),
                           outputs=("xi", "sigma"))
        xi = out["xi"]
        a.append(float(xi[6, 0])); s.append(out["sigma"][:, 0].cpu().numpy())
    w = np.array([1, 2, 2, 1, 2, 1])[None, :]
    assert np.linalg.norm(np.array(a) - alpha) < 1e-6
    assert np.sqrt((w * (np.array(s) - _vec6(stress)) ** 2).sum()) < 1e-6
