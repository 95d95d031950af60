#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -k "mixed or objective or fe_reference or reference_golden" > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
tail -4 gpurun_out/r1d_pytest.log
python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 22 --nsteps 20 > gpurun_out/r1d_k2_j2.jsonl 2> gpurun_out/r1d_k2.err
python benchmarks/mp_bench.py --what k2 --yield hosford:4 --log2n 22 --nsteps 20 --diag-only > gpurun_out/r1d_k2_hosford.jsonl 2>> gpurun_out/r1d_k2.err
python benchmarks/fe_bench.py --family hex8 --div 96 --variants K3,MIX --steps 10 > gpurun_out/r1d_fe_hex8.jsonl 2> gpurun_out/r1d_fe.err
python benchmarks/fe_bench.py --family tet4 --div 100 --variants K3,MIX --steps 10 > gpurun_out/r1d_fe_tet4.jsonl 2>> gpurun_out/r1d_fe.err
cat gpurun_out/r1d_k2_j2.jsonl gpurun_out/r1d_k2_hosford.jsonl gpurun_out/r1d_fe_hex8.jsonl gpurun_out/r1d_fe_tet4.jsonl | cut -c1-900
K2="python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 20 --nsteps 20 --steps 1 --warmup 1"
$K2 > gpurun_out/plain_k2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mp_sens -c 2 -f -o gpurun_out/prof_r1_k2_j2 $K2 > gpurun_out/ncu_k2.log 2>&1
MX="python benchmarks/fe_bench.py --family hex8 --div 48 --variants MIX --steps 1 --warmup 1"
$MX > gpurun_out/plain_mx.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fe_mixed -c 1 -f -o gpurun_out/prof_r1_mixed_hex8 $MX > gpurun_out/ncu_mx.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
