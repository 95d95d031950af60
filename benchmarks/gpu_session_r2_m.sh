#!/bin/bash
# Round-2 session M: block-level hand-off kernel (mp_update_cta.cu): parity vs the one-pass kernels, A/B timing, DRAM traffic.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -x -q -k "streaming" > gpurun_out/r2m_pytest_mp.log 2>&1; tail -5 gpurun_out/r2m_pytest_mp.log
rm -f gpurun_out/r2m_k1.jsonl
run() { timeout 300 python benchmarks/mp_bench.py --what k1 --log2n 23 --steps 5 "$@" >> gpurun_out/r2m_k1.jsonl 2>> gpurun_out/r2m_k1.err; }
run --yield hosford:4 --tag base
run --yield hosford:4 --cta --defer 0 --tag cta_k0
run --yield hosford:4 --cta --defer 1 --tag cta_k1
run --yield hosford:100 --tag base
for k in 0 1 2 3; do run --yield hosford:100 --cta --defer $k --tag cta_k$k; done
run --yield hill --tag base
for k in 0 1 2 4; do run --yield hill --cta --defer $k --tag cta_k$k; done
run --yield J2 --generic --tag base
run --yield J2 --generic --cta --defer 0 --tag cta_k0
python - <<'PY'
import json
for l in open('gpurun_out/r2m_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['checksum'][:3])
PY
tail -3 gpurun_out/r2m_k1.err
for y in hosford:4 hosford:100; do
  tag=$(echo $y | tr ':' '_'); k=0; [ "$y" = "hosford:100" ] && k=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update_cta --launch-skip 4 --launch-count 1 \
     -o gpurun_out/r2m_cta_$tag -f python benchmarks/mp_bench.py --what k1 --yield $y --log2n 21 --steps 3 --cta --defer $k > gpurun_out/r2m_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2m_ncu_$tag.log
done
