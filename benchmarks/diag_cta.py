"""Diagnostic: block-level hand-off kernel vs one-pass kernel, entry-by-entry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmad_b200 import NewtonSettings, Parameters, active_param_ids, material_from_values, mp
from tests.helpers import param_tree, random_strains
dev = torch.device("cuda:0")
rng = np.random.default_rng(17)
values, act, tr = param_tree("hosford", ("voce", "linear"), a=4.0, active=("E", "nu", "D", "S", "Y", "K"))
P = Parameters(values, act, tr); mat = material_from_values(values); pid = active_param_ids(P)
kw = dict(max_iters=20, abs_tol=1e-12, rel_tol=1e-12)
for n, scale in ((4099, 1.2e-3), (4096, 1.2e-3), (512, 1.2e-3), (4099, 0.6e-3), (4099, 3e-3)):
    e = random_strains(rng, n, scale=scale, diag_only=True)
    xi = torch.zeros((7, n), dtype=torch.float64, device=dev)
    ed = torch.from_numpy(e).to(dev)
    keys = ("xi", "iters", "flags", "cnorm", "sigma", "dsig_deps", "dC_dp")
    a = mp.mp_update(mat, NewtonSettings(cta=True, defer_after=0, **kw), pid, xi, ed, outputs=keys)
    b = mp.mp_update(mat, NewtonSettings(one_pass=True, defer_after=0, **kw), pid, xi, ed, outputs=keys)
    torch.cuda.synchronize()
    pl = float(((b["flags"] & 2) != 0).double().mean())
    line = [f"n={n} scale={scale} plastic={pl:.2f}"]
    for k in keys:
        da = (a[k].double() - b[k].double()).abs()
        bad = (da > 0)
        cols = bad.any(dim=0) if bad.dim() == 2 else bad
        idx = torch.nonzero(cols).reshape(-1)[:8].tolist()
        line.append(f"{k}: {int(cols.sum())} pts differ, max {float(da.max()):.3e}, first {idx}")
    print(" | ".join(line))
    if int((a["iters"] != b["iters"]).sum()):
        j = torch.nonzero(a["iters"] != b["iters"]).reshape(-1)[:8]
        print("   iters cta", a["iters"][j].tolist(), "one-pass", b["iters"][j].tolist(), "tile-local", (j % 512).tolist())
