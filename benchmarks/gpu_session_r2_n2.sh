#!/bin/bash
# N = 2 check of the bench contract (torchrun launch as the driver does), both arms.
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err ); tail -n 3 gpurun_out/r2z_bench_n2.err; cut -c1-300 gpurun_out/r2z_bench_n2.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2z_ref_n2.json 2> gpurun_out/r2z_ref_n2.err ); cut -c1-200 gpurun_out/r2z_ref_n2.json
