#!/bin/bash
# Round-2 session J: hex8 K3 with / without the L2 prefetch of a later block's inputs; Hosford after the
# division / sqrt special-case removals.
mkdir -p gpurun_out
rm -f gpurun_out/r2j_fe.jsonl gpurun_out/r2j_k1.jsonl
for pf in 0 1 2 3 6; do
  CMADX_HEX8_PREFETCH=$pf timeout 600 python benchmarks/fe_bench.py --family hex8 --div 128 --steps 10 --variants K3 >> gpurun_out/r2j_fe.jsonl 2>> gpurun_out/r2j_fe.err
done
for y in hosford:4 hosford:100 hill; do
  timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 >> gpurun_out/r2j_k1.jsonl 2>> gpurun_out/r2j_k1.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r2j_fe.jsonl', 'gpurun_out/r2j_k1.jsonl'):
    for l in open(f):
        d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('tag', 'family', 'variant', 'name', 'kernel', 'yield', 'ms_per_step', 'ms', 'frac_hbm', 'what', 'elements')})
PY
tail -3 gpurun_out/r2j_fe.err gpurun_out/r2j_k1.err
