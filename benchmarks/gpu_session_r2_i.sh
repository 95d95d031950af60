#!/bin/bash
# Round-2 session I: full GPU suite after the block-golden / plugin / lock-step changes, then the default bench.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_pytest.log 2>&1; tail -6 gpurun_out/r2i_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ); tail -3 gpurun_out/r2i_bench.err; cut -c1-300 gpurun_out/r2i_bench.json
