#!/bin/bash
# Round-2 session K = I + J: full GPU suite, hex8 prefetch A/B, generic K1 timings, default bench.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2k_pytest.log 2>&1; tail -6 gpurun_out/r2k_pytest.log
bash benchmarks/gpu_session_r2_j.sh
( time timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err ); tail -3 gpurun_out/r2k_bench.err; cut -c1-300 gpurun_out/r2k_bench.json
