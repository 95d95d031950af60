"""The raw AD products of the reference's `Model` object (cmad/models/model.py:121-160) - dC/dU,
dC/dU_prev, dcauchy/dxi, dcauchy/dxi_prev, dcauchy/dparams, and dcauchy/dU - against the
reference's own run (tests/golden/ref_model_partials.npz, written by
`make_reference_golden.py --only partials`: `model._jacobian[DU | DU_PREV]`, `model.dcauchy[...]`
at converged states of J2 / rotated Hill / Hosford / rotated Yld2004-18p points).

CPU: the torch-AD oracle.  GPU: `cmadx_mp_model_partials` (closed forms).  1e-10 relative."""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import const_like, material
from tests.helpers import rel_err
from tests.test_reference_golden import _ROW9, _sym_cols

D = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_partials.npz"))
KINDS = ["J2", "hill_rot", "hosford", "barlat_rot"]


def test_fixture_zero_blocks_and_plastic_content():
    """dC/dU_prev and dcauchy/dxi_prev vanish identically for SmallElasticPlastic (the C-ABI has no
    output for them), and the plastic branch is exercised."""
    for kind in KINDS:
        assert np.abs(D[f"{kind}.dC_dU_prev"]).max() == 0.0 and np.abs(D[f"{kind}.dsig_dxi_prev"]).max() == 0.0
        assert (D[f"{kind}.flags"] > 0).sum() >= 3 and (D[f"{kind}.flags"] == 0).sum() >= 1
        assert np.abs(D[f"{kind}.dC_dU"][D[f"{kind}.flags"] > 0]).max() > 0


@pytest.mark.parametrize("kind", KINDS)
def test_torch_oracle_partials_vs_reference(kind):
    import torch
    from oracle import cmad_oracle as co
    tv = co.to_torch_tree(material(kind))
    spec = co.ModelSpec()
    for i in range(0, D[f"{kind}.xi"].shape[0], 3):
        x, xp = torch.from_numpy(D[f"{kind}.xi"][i].copy()), torch.from_numpy(D[f"{kind}.xi_prev"][i].copy())
        gu = torch.from_numpy(D[f"{kind}.grad_u"][i].reshape(3, 3).copy())
        assert rel_err(co.dC_dgrad_u(x, xp, tv, gu, gu, spec).reshape(7, 9).numpy(), D[f"{kind}.dC_dU"][i]) < 1e-10
        assert rel_err(co.dcauchy_dxi(x, xp, tv, gu, gu, spec).reshape(9, 7).numpy(), D[f"{kind}.dsig_dxi"][i]) < 1e-10
        assert rel_err(co.dcauchy_dgrad_u(x, xp, tv, gu, gu, spec).reshape(9, 9).numpy(), D[f"{kind}.dsig_dU"][i]) < 1e-10
        dp = co.dcauchy_dparams(x, xp, tv, gu, gu, spec)
        el = np.stack([dp["elastic"]["E"].reshape(9).numpy(), dp["elastic"]["nu"].reshape(9).numpy()], axis=1)
        assert rel_err(el, D[f"{kind}.dsig_dp"][i][:, :2]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("comps", [9, 6])
def test_cuda_model_partials_vs_reference(cuda_device, kind, comps):
    import torch
    from cmad_b200 import active_param_ids, material_from_values, mp
    values = material(kind)
    act = const_like(values, False)
    act["elastic"] = {k: True for k in values["elastic"]}
    act["plastic"]["flow stress"]["initial yield"]["Y"] = True            # a leaf cauchy does not see: zero column
    P = Parameters(values, act, const_like(values, None))
    aidx = np.asarray(P.active_idx)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(cuda_device)  # noqa: E731
    gu = D[f"{kind}.grad_u"]
    sym6 = np.stack([gu[:, 0], 0.5 * (gu[:, 1] + gu[:, 3]), 0.5 * (gu[:, 2] + gu[:, 6]),
                     gu[:, 4], 0.5 * (gu[:, 5] + gu[:, 7]), gu[:, 8]], axis=1)
    strain = gu if comps == 9 else sym6
    out = mp.model_partials(material_from_values(values), active_param_ids(P), t(D[f"{kind}.xi"]),
                            t(D[f"{kind}.xi_prev"]), t(strain))
    torch.cuda.synchronize()
    n = gu.shape[0]
    o = {k: v.cpu().numpy().T for k, v in out.items()}
    assert rel_err(o["dC_deps"].reshape(n, 7, 6), _sym_cols(D[f"{kind}.dC_dU"])) < 1e-10
    assert rel_err(o["dsig_dxi"].reshape(n, 6, 7), D[f"{kind}.dsig_dxi"][:, _ROW9, :]) < 1e-10
    assert rel_err(o["dsig_deps"].reshape(n, 6, 6), _sym_cols(D[f"{kind}.dsig_dU"])[:, _ROW9, :]) < 1e-10
    ref_dp = D[f"{kind}.dsig_dp"][:, _ROW9, :][:, :, aidx]
    assert rel_err(o["dsig_dp"].reshape(n, 6, len(aidx)), ref_dp) < 1e-10
    assert np.abs(ref_dp[:, :, -1]).max() == 0.0 and np.abs(o["dsig_dp"].reshape(n, 6, len(aidx))[:, :, -1]).max() == 0.0
