#!/bin/bash
# Session-4 final verification: whole GPU suite, smoke(), bench (both arms), ncu launch list of the bench
# command, FE numbers at config-4 sizes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s4f_pytest_full.log 2>&1; tail -2 gpurun_out/s4f_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s4f_smoke.log 2>&1; tail -1 gpurun_out/s4f_smoke.log
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/s4f_ref.json 2> gpurun_out/s4f_ref.err
python bench.py > gpurun_out/s4f_bench.json 2> gpurun_out/s4f_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/s4f_bench.json')); r = json.load(open('gpurun_out/s4f_ref.json'))
print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'ref', r['value'], d['clocks'])
PY
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/s4f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s4f_launches.csv $B > gpurun_out/s4f_ncu_l.log 2>&1
python benchmarks/fe_bench.py --family tet4 --div 119 --variants K3,K4,K5 --steps 10 > gpurun_out/s4f_fe_tet4_j2.jsonl 2> gpurun_out/s4f_fe.err
python benchmarks/fe_bench.py --family tet4 --div 119 --yield hosford:100 --variants K3 --steps 5 > gpurun_out/s4f_fe_tet4_hosford100.jsonl 2>> gpurun_out/s4f_fe.err
cut -c1-330 gpurun_out/s4f_fe_tet4_*.jsonl | cut -c150-330
