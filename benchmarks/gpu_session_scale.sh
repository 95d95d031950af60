#!/bin/bash
# Multi-GPU session (N = $1): bench.py, MP config3/calib, FE adjoint step.  All timing on the device.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err
tail -c 1500 gpurun_out/scale_bench_n$N.json
$TR --master-port 29502 benchmarks/mp_multi_gpu.py > gpurun_out/scale_mp_n$N.jsonl 2> gpurun_out/scale_mp_n$N.err
cat gpurun_out/scale_mp_n$N.jsonl
$TR --master-port 29503 benchmarks/fe_multi_gpu.py --family tet4 --div 80 > gpurun_out/scale_fe_tet4_n$N.json 2> gpurun_out/scale_fe_n$N.err
$TR --master-port 29504 benchmarks/fe_multi_gpu.py --family hex8 --div 80 > gpurun_out/scale_fe_hex8_n$N.json 2>> gpurun_out/scale_fe_n$N.err
cat gpurun_out/scale_fe_tet4_n$N.json gpurun_out/scale_fe_hex8_n$N.json
for f in gpurun_out/scale_*_n$N.err; do tail -n 2 $f; done
