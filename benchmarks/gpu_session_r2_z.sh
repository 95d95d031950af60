#!/bin/bash
# Round-2 session Z: whole GPU suite after the last additions, Barlat K1 timing (default launch bounds),
# default bench (both arms) + launch list of the bench command.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r2z_pytest.log 2>&1; tail -n 8 gpurun_out/r2z_pytest.log
rm -f gpurun_out/r2z_k1.jsonl
for y in barlat:8 barlat:18.2 hill; do timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 21 --steps 5 >> gpurun_out/r2z_k1.jsonl 2>> gpurun_out/r2z_k1.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2z_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm')
PY
( time timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err ); tail -n 2 gpurun_out/r2z_bench.err; cut -c1-400 gpurun_out/r2z_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err ); cut -c1-200 gpurun_out/r2z_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra-steps 2 > gpurun_out/r2z_ncu_bench.log 2>&1
wc -l gpurun_out/r2z_launches.csv
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
