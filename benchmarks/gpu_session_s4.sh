#!/bin/bash
# Session-4 verification: whole GPU suite, smoke(), bench (both arms), ncu launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest_full.log 2>&1; tail -2 gpurun_out/s4_pytest_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s4_smoke.log 2>&1; tail -1 gpurun_out/s4_smoke.log
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/s4_ref.json 2> gpurun_out/s4_ref.err
python bench.py > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
tail -c 900 gpurun_out/s4_bench.json; echo
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/s4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s4_launches.csv $B > gpurun_out/s4_ncu_l.log 2>&1
python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 22 --nsteps 20 > gpurun_out/s4_k2_j2.jsonl 2> gpurun_out/s4_k2.err
cut -c120-520 gpurun_out/s4_k2_j2.jsonl
