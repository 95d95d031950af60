#!/bin/bash
# Round-2 last session: whole GPU suite after the rate-form / rotated def-type Hessian kernels,
# default bench (both arms) + launch list of the bench command + smoke.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r2zz_pytest.log 2>&1; tail -n 8 gpurun_out/r2zz_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err ); tail -n 2 gpurun_out/r2zz_bench.err; cut -c1-400 gpurun_out/r2zz_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2zz_ref.json 2> gpurun_out/r2zz_ref.err ); cut -c1-200 gpurun_out/r2zz_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra-steps 2 > gpurun_out/r2zz_ncu_bench.log 2>&1
wc -l gpurun_out/r2zz_launches.csv
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
