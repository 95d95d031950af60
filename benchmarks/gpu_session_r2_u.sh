#!/bin/bash
# Round-2 session U: new rows (rotated Hessian, UniaxialCalibration, host entry points), K2 J2 timings, bench with the objective e2e.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r2u_pytest.log 2>&1; tail -n 8 gpurun_out/r2u_pytest.log
rm -f gpurun_out/r2u_k2.jsonl
timeout 600 python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 22 --nsteps 20 --steps 5 >> gpurun_out/r2u_k2.jsonl 2>> gpurun_out/r2u_k2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2u_k2.jsonl'):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('kernel', 'yield', 'ms_per_step', 'ms_min', 'frac_hbm', 'point_steps_per_s')})
PY
tail -n 3 gpurun_out/r2u_k2.err
( time timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err ); tail -n 2 gpurun_out/r2u_bench.err; cut -c1-200 gpurun_out/r2u_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2u_bench.json'))
for c in d['extra']['configs']:
    print(c['name'], c.get('ms_per_step'), c.get('frac_binding'), c.get('e2e'), c.get('error'))
PY
