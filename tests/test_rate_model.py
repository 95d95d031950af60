"""SmallRateElasticPlastic (cmad/models/small_rate_elastic_plastic.py:34-100, 250-346), the
rate form of the small-strain model (state = [cauchy(6), alpha]), FULL_3D - against golden
vectors produced by executing the reference's own source (make_reference_golden.py section F,
fixture ref_rate_model.npz): the `Model` object under the imperative `newton_solve` along a
two-leg history ((ii, ||C||), xi, Sigma, dC/dxi, dC/dxi_prev, dC/dp) and `make_newton_solve`
+ its IFT rule from the same previous states.

CPU: the torch-AD oracle's restatement (`cmad_oracle.rate_residual`) vs the fixtures.
GPU: `mp_update_rate.cu` through the C-ABI vs the fixtures (values 1e-10, derivatives 1e-9,
iteration counts exactly equal).
"""
import os

import numpy as np
import pytest

from cmad_b200 import Parameters
from tests.golden.materials import active_all_scalars, const_like, material
from tests.helpers import UP, rel_err

RT = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rate_model.npz"))
KINDS = sorted({k.split(".")[0] for k in RT.files})


@pytest.mark.parametrize("kind", KINDS)
def test_torch_oracle_vs_reference_rate_model(kind):
    import torch
    from oracle import cmad_oracle as co
    spec = co.ModelSpec(kind="small_rate_elastic_plastic")
    tv = co.to_torch_tree(material(kind))
    F = RT[f"{kind}.F"]
    x = torch.zeros(7, dtype=torch.float64)
    for t in range(1, 13):
        gu = torch.from_numpy(F[:, :, t] - np.eye(3)); gup = torch.from_numpy(F[:, :, t - 1] - np.eye(3))
        xn, info = co.newton_imperative(x, tv, gu, gup, spec)
        assert info.iters == RT[f"{kind}.iters"][t - 1], (kind, t)
        assert rel_err(xn.numpy(), RT[f"{kind}.xi"][t - 1]) < 1e-10
        assert rel_err(co.dC_dxi(xn, x, tv, gu, gup, spec).numpy(), RT[f"{kind}.dC_dxi"][t - 1]) < 1e-9
        x = xn
    assert RT[f"{kind}.xi"][:, 6].max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_vs_reference_rate_model(cuda_device, kind):
    import torch
    from cmad_b200 import NewtonSettings, active_param_ids, material_from_values, mp
    values = material(kind)
    P = Parameters(values, active_all_scalars(values), const_like(values, None))
    mat = material_from_values(values, model="small_rate_elastic_plastic")
    pid, aidx = active_param_ids(P), np.asarray(P.active_idx)
    F = RT[f"{kind}.F"]
    N = F.shape[2] - 1
    want = ("xi", "sigma", "iters", "cnorm", "dC_dxi", "dC_dxi_prev", "dC_dp", "dxi_deps", "dsig_deps")
    cols = [((3 * k + l,) if k == l else (3 * k + l, 3 * l + k)) for k, l in UP]
    xi = torch.zeros((7, 1), dtype=torch.float64, device=cuda_device)
    for t in range(1, N + 1):
        dgu = (F[:, :, t] - F[:, :, t - 1]).reshape(9, 1)                 # the kernel takes the INCREMENT
        e = torch.from_numpy(dgu.copy()).to(cuda_device)
        ot = mp.mp_update(mat, NewtonSettings(mode="traced"), pid, xi, e, outputs=want)
        assert int(ot["iters"][0]) == RT[f"{kind}.traced_iters"][t - 1], (kind, t, "traced count")
        assert rel_err(ot["xi"][:, 0].cpu().numpy(), RT[f"{kind}.traced_xi"][t - 1]) < 1e-10
        dx = np.stack([sum(RT[f"{kind}.dxi_dgradu"][t - 1][:, c] for c in cc) for cc in cols], axis=-1)   # (7, 6)
        assert rel_err(ot["dxi_deps"][:, 0].cpu().numpy().reshape(7, 6), dx) < 1e-8, (kind, t, "dxi_deps")
        assert rel_err(ot["dsig_deps"][:, 0].cpu().numpy().reshape(6, 6), dx[:6]) < 1e-8
        o = mp.mp_update(mat, NewtonSettings(mode="imperative"), pid, xi, e, outputs=want)
        assert int(o["iters"][0]) == RT[f"{kind}.iters"][t - 1], (kind, t)
        assert abs(float(o["cnorm"][0]) - RT[f"{kind}.cnorm"][t - 1]) < 1e-11
        assert rel_err(o["xi"][:, 0].cpu().numpy(), RT[f"{kind}.xi"][t - 1]) < 1e-10, (kind, t)
        assert rel_err(o["sigma"][:, 0].cpu().numpy(), RT[f"{kind}.sigma"][t - 1]) < 1e-10
        assert rel_err(o["dC_dxi"][:, 0].cpu().numpy().reshape(7, 7), RT[f"{kind}.dC_dxi"][t - 1]) < 1e-9
        assert rel_err(o["dC_dxi_prev"][:, 0].cpu().numpy().reshape(7, 7), RT[f"{kind}.dC_dxi_prev"][t - 1]) < 1e-9
        assert rel_err(o["dC_dp"][:, 0].cpu().numpy().reshape(7, len(aidx)),
                       RT[f"{kind}.dC_dp"][t - 1][:, aidx]) < 1e-9, (kind, t, "dC_dp")
        xi = o["xi"]
