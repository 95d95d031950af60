#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/defer_pytest.log 2>&1; tail -6 gpurun_out/defer_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 benchmarks/mp_multi_gpu.py 2>/dev/null > gpurun_out/defer_mp.jsonl; cut -c1-300 gpurun_out/defer_mp.jsonl
python benchmarks/mp_bench.py --what k2 --yield J2 --log2n 22 --nsteps 20 2>/dev/null > gpurun_out/defer_k2.jsonl; cut -c230-420 gpurun_out/defer_k2.jsonl
python benchmarks/mp_bench.py --what k1 --yield hill --log2n 23 2>/dev/null > gpurun_out/defer_k1_hill.jsonl; cut -c1-420 gpurun_out/defer_k1_hill.jsonl
python benchmarks/fe_bench.py --family tet4 --div 119 --yield hosford:100 --variants K3 --steps 5 2>/dev/null > gpurun_out/defer_fe.jsonl
python benchmarks/fe_bench.py --family tet4 --div 119 --yield hosford:4 --variants K3 --steps 5 2>/dev/null >> gpurun_out/defer_fe.jsonl
python benchmarks/fe_bench.py --family hex8 --div 96 --yield hosford:4 --variants K3 --steps 5 2>/dev/null >> gpurun_out/defer_fe.jsonl
cut -c1-420 gpurun_out/defer_fe.jsonl
