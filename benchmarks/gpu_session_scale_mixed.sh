#!/bin/bash
# Multi-GPU session (N = $1): bench.py + the mixed u-p adjoint step (configs[4]-style) + the displacement-form step.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s4_scale_bench_n$N.json 2> gpurun_out/s4_scale_bench_n$N.err
tail -c 700 gpurun_out/s4_scale_bench_n$N.json; echo
for f in tet4 hex8; do
  $TR --master-port 29503 benchmarks/fe_multi_gpu.py --family $f --div 80 --mixed > gpurun_out/s4_scale_fe_mixed_${f}_n$N.json 2> gpurun_out/s4_scale_fe_n$N.err
  $TR --master-port 29504 benchmarks/fe_multi_gpu.py --family $f --div 80 > gpurun_out/s4_scale_fe_${f}_n$N.json 2>> gpurun_out/s4_scale_fe_n$N.err
done
cat gpurun_out/s4_scale_fe_*_n$N.json | cut -c1-520
$TR --master-port 29502 benchmarks/mp_multi_gpu.py > gpurun_out/s4_scale_mp_n$N.jsonl 2> gpurun_out/s4_scale_mp_n$N.err
cut -c1-400 gpurun_out/s4_scale_mp_n$N.jsonl
tail -n 2 gpurun_out/s4_scale_*_n$N.err
