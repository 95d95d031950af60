#!/bin/bash
# Round-2 session T: full GPU suite, K2 direct with dxi/dp in shared memory, ncu of the hex8 K3 / tet4 x 4 kernels,
# default bench (both arms) + launch list.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r2t_pytest.log 2>&1; tail -8 gpurun_out/r2t_pytest.log
rm -f gpurun_out/r2t_k2.jsonl
timeout 600 python benchmarks/mp_bench.py --what k2 --log2n 22 --nsteps 20 --steps 5 >> gpurun_out/r2t_k2.jsonl 2>> gpurun_out/r2t_k2.err
timeout 600 python benchmarks/mp_bench.py --what k2 --yield hosford:4 --log2n 21 --nsteps 20 --steps 5 >> gpurun_out/r2t_k2.jsonl 2>> gpurun_out/r2t_k2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2t_k2.jsonl'):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('kernel', 'yield', 'ms_per_step', 'ms_min', 'frac_hbm', 'point_steps_per_s')})
PY
tail -n 3 gpurun_out/r2t_k2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_hex8_kernel --launch-skip 6 --launch-count 1 \
   -o gpurun_out/r2t_hex8_k3 -f python benchmarks/fe_bench.py --family hex8 --div 128 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2t_ncu_hex8.log 2>&1; tail -n 1 gpurun_out/r2t_ncu_hex8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_tet4x4_kernel --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r2t_tet4x4 -f python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2t_ncu_tet4x4.log 2>&1; tail -n 1 gpurun_out/r2t_ncu_tet4x4.log
( time timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err ); tail -n 2 gpurun_out/r2t_bench.err; cut -c1-300 gpurun_out/r2t_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2t_ref.json 2> gpurun_out/r2t_ref.err ); cut -c1-200 gpurun_out/r2t_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra-steps 2 > gpurun_out/r2t_ncu_bench.log 2>&1
wc -l gpurun_out/r2t_launches.csv
