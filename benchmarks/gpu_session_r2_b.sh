#!/bin/bash
# Round-2 session B: streaming (lane-refill) generic K1 vs the one-pass kernels: parity tests, A/B timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -x -q > gpurun_out/r2b_pytest_mp.log 2>&1; tail -5 gpurun_out/r2b_pytest_mp.log
for y in hosford:4 hosford:100 hill J2; do
  tag=$(echo $y | tr ':' '_')
  g=""; [ "$y" = "J2" ] && g="--generic"
  for v in "" "--one-pass"; do
    timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 $g $v >> gpurun_out/r2b_k1_ab.jsonl 2>> gpurun_out/r2b_k1_ab.err
  done
done
timeout 300 python benchmarks/mp_bench.py --what k1 --yield hosford:100 --log2n 23 --steps 5 --max-iters 500 --ls-evals 100 >> gpurun_out/r2b_k1_ab.jsonl 2>> gpurun_out/r2b_k1_ab.err
timeout 300 python benchmarks/mp_bench.py --what k1 --yield hosford:100 --log2n 23 --steps 5 --max-iters 500 --ls-evals 100 --one-pass >> gpurun_out/r2b_k1_ab.jsonl 2>> gpurun_out/r2b_k1_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r2b_k1_ab.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('newton'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['mean_newton_iters'])
PY
tail -3 gpurun_out/r2b_k1_ab.err
