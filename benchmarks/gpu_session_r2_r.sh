#!/bin/bash
# Round-2 session R: full GPU suite (no -x), tet4 x 4 factored kernel, Hosford after the cheaper outer root,
# ncu of the hex8 K3 kernel (the right instantiation this time).
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/r2r_pytest.log 2>&1; tail -12 gpurun_out/r2r_pytest.log
rm -f gpurun_out/r2r_fe.jsonl gpurun_out/r2r_k1.jsonl
timeout 600 python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 10 --variants K3,K4,MIX >> gpurun_out/r2r_fe.jsonl 2>> gpurun_out/r2r_fe.err
CMADX_TET4X4_GENERAL=1 timeout 600 python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 10 --variants K3 >> gpurun_out/r2r_fe.jsonl 2>> gpurun_out/r2r_fe.err
for y in hosford:4 hosford:100 hill; do
  timeout 300 python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 >> gpurun_out/r2r_k1.jsonl 2>> gpurun_out/r2r_k1.err
done
python - <<'PY'
import json
for f in ('gpurun_out/r2r_fe.jsonl', 'gpurun_out/r2r_k1.jsonl'):
    for l in open(f):
        d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('family', 'kernel', 'n_elems', 'n_ip', 'yield', 'solver', 'ms_per_step', 'ms_min', 'frac_hbm')})
PY
tail -3 gpurun_out/r2r_fe.err gpurun_out/r2r_k1.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fe_hex8_kernel<0' --launch-skip 2 --launch-count 1 \
   -o gpurun_out/r2r_hex8_k3 -f python benchmarks/fe_bench.py --family hex8 --div 128 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2r_ncu_hex8.log 2>&1; tail -1 gpurun_out/r2r_ncu_hex8.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fe_tet4x4_kernel<0' --launch-skip 2 --launch-count 1 \
   -o gpurun_out/r2r_tet4x4 -f python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2r_ncu_tet4x4.log 2>&1; tail -1 gpurun_out/r2r_ncu_tet4x4.log
