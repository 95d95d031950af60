#!/bin/bash
# Round-2 session P: FE kernels after the bulk-store (hex8 K_e), 2x12 pressure tiling (hex8 mixed) and
# reduce-scatter (tet4 x 4) changes: FE GPU tests, memcheck of a small hex8 block, A/B timings.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_fe.py tests/test_fe_block_reference_golden.py tests/test_fe_reference_golden.py tests/test_fe_driver.py tests/test_legacy_line_search.py -m gpu -x -q ) > gpurun_out/r2p_pytest.log 2>&1; tail -6 gpurun_out/r2p_pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fe.py -m gpu -x -q -k "hex8 or mixed" > gpurun_out/r2p_memcheck.log 2>&1; tail -4 gpurun_out/r2p_memcheck.log
rm -f gpurun_out/r2p_fe.jsonl
timeout 600 python benchmarks/fe_bench.py --family hex8 --div 128 --steps 10 --variants K3,K4 >> gpurun_out/r2p_fe.jsonl 2>> gpurun_out/r2p_fe.err
timeout 600 python benchmarks/fe_bench.py --family hex8 --div 96 --steps 10 --variants MIX >> gpurun_out/r2p_fe.jsonl 2>> gpurun_out/r2p_fe.err
CMADX_PRESSURE_ROW_KERNEL=1 timeout 600 python benchmarks/fe_bench.py --family hex8 --div 96 --steps 10 --variants MIX >> gpurun_out/r2p_fe.jsonl 2>> gpurun_out/r2p_fe.err
timeout 600 python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 10 --variants K3,MIX >> gpurun_out/r2p_fe.jsonl 2>> gpurun_out/r2p_fe.err
CMADX_TET4X4_FLAT=1 timeout 600 python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 10 --variants K3,MIX >> gpurun_out/r2p_fe.jsonl 2>> gpurun_out/r2p_fe.err
python - <<'PY'
import json
for l in open('gpurun_out/r2p_fe.jsonl'):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ('family', 'kernel', 'n_elems', 'ms_per_step', 'ms_min', 'frac_hbm')})
PY
tail -3 gpurun_out/r2p_fe.err
