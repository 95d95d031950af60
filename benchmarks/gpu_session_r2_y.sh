#!/bin/bash
# Round-2 session Y: rate model under rotated axes / in the mixed u-p form, def-type kernels after the
# point-interface refactor; Barlat K1 timing; then the whole suite.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_rate_rot_mixed.py tests/test_rate_model.py tests/test_def_types.py tests/test_hessian.py -m gpu -q ) > gpurun_out/r2y_new.log 2>&1; tail -n 40 gpurun_out/r2y_new.log
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2y_pytest.log 2>&1; tail -n 6 gpurun_out/r2y_pytest.log
rm -f gpurun_out/r2y_k1.jsonl
for y in barlat:8 barlat:18.2 hill; do timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 21 --steps 5 >> gpurun_out/r2y_k1.jsonl 2>> gpurun_out/r2y_k1.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2y_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', 'plastic', round(d['plastic_fraction'],3), 'iters', round(d['mean_newton_iters'],2))
PY
tail -n 3 gpurun_out/r2y_k1.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
