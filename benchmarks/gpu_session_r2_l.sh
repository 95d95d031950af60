#!/bin/bash
# Round-2 session L: 7x7 generic K1 at 3 blocks / SM (168 registers); DRAM traffic of the two passes of Hosford a = 100;
# deck-driven GPU tests.
mkdir -p gpurun_out
rm -f gpurun_out/r2l_k1.jsonl
for y in hill J2; do
  g=""; [ "$y" = "J2" ] && g="--generic"
  for o in 0 1; do
    CMADX_SEP_OCC3=$o timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 $g --tag occ3_$o >> gpurun_out/r2l_k1.jsonl 2>> gpurun_out/r2l_k1.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2l_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['checksum'][:2])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:mp_update --launch-skip 8 --launch-count 4 --csv --log-file gpurun_out/r2l_a100_passes.csv python benchmarks/mp_bench.py --what k1 --yield hosford:100 --log2n 23 --steps 3 > gpurun_out/r2l_ncu.log 2>&1
cat gpurun_out/r2l_a100_passes.csv | cut -d, -f5,13- | tail -20
timeout 900 python -m pytest tests/test_deck.py -m gpu -x -q > gpurun_out/r2l_pytest_deck.log 2>&1; tail -5 gpurun_out/r2l_pytest_deck.log
