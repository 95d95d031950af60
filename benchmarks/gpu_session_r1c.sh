#!/bin/bash
# One GPU session: full -m gpu suite, bench, mixed FE bench lines.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c_pytest.log
tail -5 gpurun_out/r1c_pytest.log
python benchmarks/fe_bench.py --family hex8 --div 96 --variants K3,MIX,K4 --steps 10 > gpurun_out/r1c_fe_hex8.jsonl 2> gpurun_out/r1c_fe_hex8.err
python benchmarks/fe_bench.py --family tet4 --div 100 --variants K3,MIX,K4 --steps 10 > gpurun_out/r1c_fe_tet4.jsonl 2> gpurun_out/r1c_fe_tet4.err
cat gpurun_out/r1c_fe_hex8.jsonl gpurun_out/r1c_fe_tet4.jsonl
