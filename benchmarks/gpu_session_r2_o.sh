#!/bin/bash
# Round-2 session O: full GPU suite, default bench (both arms), launch list of the bench command, A/B of the final defaults.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2o_pytest.log 2>&1; tail -6 gpurun_out/r2o_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err ); tail -2 gpurun_out/r2o_bench.err; cut -c1-200 gpurun_out/r2o_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2o_ref.json 2> gpurun_out/r2o_ref.err ); cut -c1-200 gpurun_out/r2o_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra-steps 2 > gpurun_out/r2o_ncu.log 2>&1
wc -l gpurun_out/r2o_launches.csv
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
from cmad_b200 import NewtonSettings, Parameters, material_from_values, mp
from oracle import oracle_c as oc
from tests.helpers import param_tree, random_strains
rng = np.random.default_rng(6)
values, act, tr = param_tree("hosford", ("voce",), a=100.0, elastic={"E": 1000.0, "nu": 0.25}, active=())
values["plastic"]["flow stress"]["initial yield"]["Y"] = 2.0
values["plastic"]["flow stress"]["hardening"]["voce"] = {"S": 10.0, "D": 2.0}
kw = dict(max_iters=500, abs_tol=1e-12, rel_tol=1e-12)
n = 2048
e = random_strains(rng, n, scale=2e-3, diag_only=True)
out = mp.mp_update(material_from_values(values), NewtonSettings(mode="traced", ls_max_evals=100, **kw), [], torch.zeros((7, n), dtype=torch.float64, device="cuda:0"), torch.from_numpy(e).cuda(), outputs=("xi", "iters", "flags", "cnorm"))
ref = oc.mp_update(oc.describe(values, [], newton_mode="traced", ls_max_evals=100, **kw), np.zeros((7, n)), e, want=("xi", "iters", "flags", "cnorm"))
it = out["iters"].cpu().numpy()
print("a=100 random batch: counts equal on", float(np.mean(it == ref["iters"])), "flags equal on", float(np.mean(out["flags"].cpu().numpy() == ref["flags"])),
      "max iters", int(ref["iters"].max()), "non-converged (oracle)", float(np.mean(ref["cnorm"] >= 1e-12)))
PY
