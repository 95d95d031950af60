#!/bin/bash
# Round-2 session F: warp-parking generic K1 (mp_update_queue.cu): parity vs the one-pass kernels, A/B timing,
# park threshold sweep, ncu capture of the Hosford a=4 kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -x -q -k "streaming" > gpurun_out/r2f_pytest_mp.log 2>&1; tail -5 gpurun_out/r2f_pytest_mp.log
rm -f gpurun_out/r2f_k1_ab.jsonl
for y in hosford:4 hosford:100 hill; do
  for v in "--one-pass" "--queue"; do
    CMADX_DEBUG_QUEUE=1 timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 $v >> gpurun_out/r2f_k1_ab.jsonl 2>> gpurun_out/r2f_k1_ab.err
  done
done
for t in 12 16 20 28 32; do
  for y in hosford:4 hosford:100 hill; do
    CMADX_QUEUE_PARK_BELOW=$t timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 --queue --tag park$t >> gpurun_out/r2f_k1_ab.jsonl 2>> gpurun_out/r2f_k1_ab.err
  done
done
timeout 300 python benchmarks/mp_bench.py --what k1 --yield hosford:100 --log2n 23 --steps 5 --max-iters 500 --ls-evals 100 --queue >> gpurun_out/r2f_k1_ab.jsonl 2>> gpurun_out/r2f_k1_ab.err
timeout 300 python benchmarks/mp_bench.py --what k1 --yield hosford:100 --log2n 23 --steps 5 --max-iters 500 --ls-evals 100 --one-pass >> gpurun_out/r2f_k1_ab.jsonl 2>> gpurun_out/r2f_k1_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r2f_k1_ab.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), d.get('newton'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['mean_newton_iters'])
PY
grep cmadx gpurun_out/r2f_k1_ab.err | sort | uniq -c
tail -3 gpurun_out/r2f_k1_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update_queue --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r2f_queue_hosford_4 -f python benchmarks/mp_bench.py --what k1 --yield hosford:4 --log2n 21 --steps 3 --queue > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
