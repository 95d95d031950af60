#!/bin/bash
# Round-2 session Q: rate-model K2 / FE blocks vs the reference's own run, FE suite after the kernel changes,
# ncu captures of the hex8 K3 (bulk store), the 2x12 pressure kernel and the tet4 x 4 kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_rate_model.py tests/test_gpu_fe.py tests/test_fe_block_reference_golden.py tests/test_fe_reference_golden.py tests/test_fe_driver.py tests/test_fe_qoi.py tests/test_gpu_objectives.py tests/test_reference_golden.py tests/test_legacy_line_search.py -m gpu -q ) > gpurun_out/r2q_pytest.log 2>&1; tail -8 gpurun_out/r2q_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_hex8_kernel --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r2q_hex8_k3 -f python benchmarks/fe_bench.py --family hex8 --div 128 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2q_ncu_hex8.log 2>&1; tail -1 gpurun_out/r2q_ncu_hex8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_mixed_pressure --launch-skip 1 --launch-count 1 \
   -o gpurun_out/r2q_pressure -f python benchmarks/fe_bench.py --family hex8 --div 96 --steps 2 --warmup 1 --variants MIX > gpurun_out/r2q_ncu_pressure.log 2>&1; tail -1 gpurun_out/r2q_ncu_pressure.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_tet4x4 --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r2q_tet4x4 -f python benchmarks/fe_bench.py --family tet4 --div 80 --volume-degree 2 --steps 2 --warmup 1 --variants K3 > gpurun_out/r2q_ncu_tet4x4.log 2>&1; tail -1 gpurun_out/r2q_ncu_tet4x4.log
ls -la gpurun_out/r2q_*.ncu-rep
