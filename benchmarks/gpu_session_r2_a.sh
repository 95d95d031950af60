#!/bin/bash
# Round-2 session A: ncu --set full captures of the generic-Newton K1 kernels (Hosford a=4 / a=100, Hill)
# after a plain run of each command, to decide what binds them.
mkdir -p gpurun_out
for y in hosford:4 hosford:100 hill; do
  tag=$(echo $y | tr ':' '_')
  python benchmarks/mp_bench.py --what k1 --yield $y --log2n 23 --steps 5 > gpurun_out/r2a_k1_$tag.jsonl 2> gpurun_out/r2a_k1_$tag.err || continue
  cut -c1-400 gpurun_out/r2a_k1_$tag.jsonl
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update --launch-skip 6 --launch-count 2 \
     -o gpurun_out/r2a_k1_$tag -f python benchmarks/mp_bench.py --what k1 --yield $y --log2n 21 --steps 3 > gpurun_out/r2a_ncu_$tag.log 2>&1
  tail -2 gpurun_out/r2a_ncu_$tag.log
done
ls -la gpurun_out/*.ncu-rep
