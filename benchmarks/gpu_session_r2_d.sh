#!/bin/bash
# Round-2 session D: streaming kernel after the carveout fix: parity, A/B timing, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -x -q > gpurun_out/r2d_pytest_mp.log 2>&1; tail -3 gpurun_out/r2d_pytest_mp.log
timeout 900 python -m pytest tests/test_reference_golden.py tests/test_def_types.py tests/test_rate_model.py tests/test_gpu_objectives.py -m gpu -x -q > gpurun_out/r2d_pytest_ref.log 2>&1; tail -3 gpurun_out/r2d_pytest_ref.log
rm -f gpurun_out/r2d_k1_ab.jsonl
for y in hosford:4 hosford:100 hill; do
  for v in "" "--one-pass"; do
    CMADX_DEBUG_STREAM=1 timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 $v >> gpurun_out/r2d_k1_ab.jsonl 2>> gpurun_out/r2d_k1_ab.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2d_k1_ab.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('newton'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['mean_newton_iters'])
PY
grep cmadx gpurun_out/r2d_k1_ab.err | sort | uniq -c
for y in hosford:4 hill; do
  tag=$(echo $y | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update_stream --launch-skip 4 --launch-count 1 \
     -o gpurun_out/r2d_stream_$tag -f python benchmarks/mp_bench.py --what k1 --yield $y --log2n 21 --steps 3 > gpurun_out/r2d_ncu_$tag.log 2>&1
  tail -2 gpurun_out/r2d_ncu_$tag.log
done
