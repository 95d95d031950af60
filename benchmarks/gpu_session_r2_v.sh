#!/bin/bash
# Round-2 session V: block hand-off kernel with lane refill in the hard phase: parity + timings.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_mp_update.py tests/test_reference_golden.py -m gpu -q ) > gpurun_out/r2v_pytest.log 2>&1; tail -n 4 gpurun_out/r2v_pytest.log
rm -f gpurun_out/r2v_k1.jsonl
run() { timeout 300 python benchmarks/mp_bench.py --what k1 --log2n 23 --steps 5 "$@" >> gpurun_out/r2v_k1.jsonl 2>> gpurun_out/r2v_k1.err; }
run --yield hosford:100 --tag default
for k in 1 2 3; do run --yield hosford:100 --cta --defer $k --tag cta_k$k; done
run --yield hosford:4 --tag default
run --yield hosford:4 --cta --defer 0 --tag cta_k0
run --yield hosford:4 --cta --defer 1 --tag cta_k1
run --yield hill --tag default
for k in 0 1 2; do run --yield hill --cta --defer $k --tag cta_k$k; done
run --yield J2 --generic --tag default
run --yield J2 --generic --cta --defer 1 --tag cta_k1
python - <<'PY'
import json
for l in open('gpurun_out/r2v_k1.jsonl'):
    d = json.loads(l); print(d['yield'], d['solver'], d.get('tag'), round(d['ms_per_step'],3), 'ms', round(d['frac_hbm'],3), 'hbm')
PY
tail -n 3 gpurun_out/r2v_k1.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
