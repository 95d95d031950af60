#!/bin/bash
# Round-2 session S: which change broke the bit-equality of the generic K1 kernels (Hosford, traced)?  A/B of the outer root.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -q -k "streaming and hosford and traced" ) > gpurun_out/r2s_fast_root.log 2>&1; tail -3 gpurun_out/r2s_fast_root.log
( CMADX_HOSFORD_LIBM_ROOT=1 timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -q -k "streaming and hosford and traced" ) > gpurun_out/r2s_libm_root.log 2>&1; tail -3 gpurun_out/r2s_libm_root.log
( timeout 900 python -m pytest tests/test_cmad_plugin.py tests/test_deck.py -m gpu -q ) > gpurun_out/r2s_plugin.log 2>&1; tail -3 gpurun_out/r2s_plugin.log
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp
from tests.helpers import param_tree, random_strains
rng = np.random.default_rng(17)
for a in (4.0, 100.0):
    values, act, tr = param_tree("hosford", ("voce", "linear"), a=a, active=("E", "nu", "D", "S", "Y", "K"))
    mat = material_from_values(values); pid = active_param_ids(Parameters(values, act, tr))
    kw = dict(max_iters=40, abs_tol=1e-12, rel_tol=1e-12, ls_max_evals=8) if a == 100.0 else dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
    n = 70001
    xi = torch.zeros((7, n), dtype=torch.float64, device="cuda:0")
    e = np.zeros((6, n))
    for s in range(3):
        e = e * 1.3 + random_strains(rng, n, scale=1.2e-3 / (1 + s), diag_only=True)
        ed = torch.from_numpy(e).cuda()
        outs = {}
        for name, nw in (("one_pass", NewtonSettings(mode="traced", one_pass=True, defer_after=0, **kw)),
                         ("default", NewtonSettings(mode="traced", defer_after=0, **kw)),
                         ("cta", NewtonSettings(mode="traced", cta=True, defer_after=0, **kw)),
                         ("stream", NewtonSettings(mode="traced", stream=True, **kw))):
            outs[name] = mp.mp_update(mat, nw, pid, xi, ed, outputs=("xi", "iters", "flags", "cnorm"))
        b = outs["one_pass"]
        for name, o in outs.items():
            d = (o["xi"] - b["xi"]).abs().max().item() / b["xi"].abs().max().item()
            print(f"a={a} step {s} {name}: max rel xi diff {d:.3e}, iters equal {bool(torch.equal(o['iters'], b['iters']))}, "
                  f"flags equal {bool(torch.equal(o['flags'], b['flags']))}, points differing {(o['xi'] != b['xi']).any(0).sum().item()}")
        xi = b["xi"]
PY
