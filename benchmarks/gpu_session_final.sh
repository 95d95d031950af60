#!/bin/bash
# Round-end measurement refresh: bench (both arms), ncu launch list of the same command, one
# ncu --set full capture of the headline kernel, config-4-style FE numbers.
mkdir -p gpurun_out
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/final_bench.json; echo
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/ncu_final_l.log 2>&1
$B > gpurun_out/plain_final2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mp_update_j2 -s 3 -c 2 -f -o gpurun_out/prof_r1_final_k1j2 $B > gpurun_out/ncu_final_f.log 2>&1
python benchmarks/fe_bench.py --family tet4 --div 119 --yield hosford:100 --variants K3,K4 --steps 5 > gpurun_out/final_fe_tet4_hosford100.jsonl 2> gpurun_out/final_fe.err
python benchmarks/fe_bench.py --family tet4 --div 119 --yield hosford:4 --variants K3 --steps 5 >> gpurun_out/final_fe_tet4_hosford100.jsonl 2>> gpurun_out/final_fe.err
python benchmarks/fe_bench.py --family tet4 --div 119 --variants K3,K4,K5 --steps 10 > gpurun_out/final_fe_tet4_j2.jsonl 2>> gpurun_out/final_fe.err
python benchmarks/fe_bench.py --family hex8 --div 128 --variants K3,K4 --steps 10 > gpurun_out/final_fe_hex8_j2.jsonl 2>> gpurun_out/final_fe.err
cut -c1-420 gpurun_out/final_fe_*.jsonl
python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 22 --nsteps 20 > gpurun_out/final_k2_j2.jsonl 2> gpurun_out/final_k2.err
cut -c120-520 gpurun_out/final_k2_j2.jsonl
ls -la gpurun_out/final_launches.csv gpurun_out/prof_r1_final_k1j2.ncu-rep
