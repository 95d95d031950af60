#!/bin/bash
# Round-2 session H: occupancy variants of the reduced Hosford kernel (register budget vs resident warps)
mkdir -p gpurun_out
rm -f gpurun_out/r2h_k1.jsonl
for y in hosford:4 hosford:100; do
  for b in 0 5 6 3; do
    CMADX_MINB=$b timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 --tag minb$b >> gpurun_out/r2h_k1.jsonl 2>> gpurun_out/r2h_k1.err
  done
  CMADX_LOCKSTEP_BLOCK=256 timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 --tag lockstep256 >> gpurun_out/r2h_k1.jsonl 2>> gpurun_out/r2h_k1.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2h_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['checksum'][:2])
PY
tail -3 gpurun_out/r2h_k1.err
