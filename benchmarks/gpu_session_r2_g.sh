#!/bin/bash
# Round-2 session G: one-pass generic K1 with block-wide lock-step Newton voting (instruction-cache locality experiment)
mkdir -p gpurun_out
rm -f gpurun_out/r2g_k1.jsonl
for y in hosford:4 hosford:100 hill; do
  for b in 0 128 256 512; do
    CMADX_LOCKSTEP_BLOCK=$b timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 --tag lockstep$b >> gpurun_out/r2g_k1.jsonl 2>> gpurun_out/r2g_k1.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2g_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['checksum'][:2])
PY
tail -3 gpurun_out/r2g_k1.err
CMADX_LOCKSTEP_BLOCK=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update_lockstep --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r2g_lockstep512_hosford_4 -f python benchmarks/mp_bench.py --what k1 --yield hosford:4 --log2n 21 --steps 3 > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
