#!/bin/bash
# Round-2 session N: block-level hand-off kernel after the record fix + warp-voted hard rounds; FP64 flop counts.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mp_update.py -m gpu -x -q -k "streaming" > gpurun_out/r2n_pytest_mp.log 2>&1; tail -4 gpurun_out/r2n_pytest_mp.log
rm -f gpurun_out/r2n_k1.jsonl
run() { timeout 300 python benchmarks/mp_bench.py --what k1 --log2n 23 --steps 5 "$@" >> gpurun_out/r2n_k1.jsonl 2>> gpurun_out/r2n_k1.err; }
run --yield hosford:4 --tag base
run --yield hosford:4 --cta --defer 0 --tag cta_k0
run --yield hosford:100 --tag base
for k in 1 2 3; do run --yield hosford:100 --cta --defer $k --tag cta_k$k; done
run --yield hosford:100 --max-iters 500 --ls-evals 100 --tag base_notch
run --yield hosford:100 --max-iters 500 --ls-evals 100 --cta --defer 2 --tag cta_k2_notch
run --yield hill --tag base
for k in 0 4; do run --yield hill --cta --defer $k --tag cta_k$k; done
run --yield J2 --generic --tag base
run --yield J2 --generic --cta --defer 0 --tag cta_k0
python - <<'PY'
import json
for l in open('gpurun_out/r2n_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm', d['checksum'][:2])
PY
tail -3 gpurun_out/r2n_k1.err
bash benchmarks/count_flops.sh
