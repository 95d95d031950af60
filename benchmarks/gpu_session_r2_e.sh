#!/bin/bash
# Round-2 session E (re-entry): full GPU suite, the default bench (both arms) with wall time,
# and the streaming generic-Newton K1 A/B against the one-pass kernels.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err ); tail -3 gpurun_out/r2e_bench.err; cut -c1-600 gpurun_out/r2e_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2e_ref.json 2> gpurun_out/r2e_ref.err ); cut -c1-300 gpurun_out/r2e_ref.json
rm -f gpurun_out/r2e_k1_ab.jsonl
for y in hosford:4 hosford:100 hill; do
  for v in "" "--one-pass"; do
    timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 $v >> gpurun_out/r2e_k1_ab.jsonl 2>> gpurun_out/r2e_k1_ab.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2e_k1_ab.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('newton'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['mean_newton_iters'])
PY
