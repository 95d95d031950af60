#!/bin/bash
# ncu --set full of the headline kernel (round-2 binary: DevMat grew by the Barlat coefficients) and of the
# Barlat K1 kernel (where does an untuned 250-double state spend its time).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_update_j2_kernel --launch-skip 3 --launch-count 1 \
   -o gpurun_out/r2z_k1_j2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra-steps 0 > gpurun_out/r2z_ncu_j2.log 2>&1; tail -n 1 gpurun_out/r2z_ncu_j2.log
ls -la gpurun_out/*.ncu-rep | tail -3
