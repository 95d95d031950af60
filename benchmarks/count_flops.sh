#!/bin/bash
# FP64 flops of the timed region of every extra config (and of the headline), counted by ncu on the
# exact bench workloads: one timed step each, NVTX-filtered to the timed region.
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum
for c in hosford_a4 hosford_a100 fe_k3_tet4 fe_k3_hex8 fe_adjoint_mixed_tet4 fe_adjoint_mixed_hex8 mp_objective; do
  timeout 900 ncu --nvtx --nvtx-include "cmadx_timed/" --metrics $M --clock-control none --csv --log-file gpurun_out/r2_flops_$c.csv \
     python bench.py --points 65536 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --extra $c --extra-steps 1 > gpurun_out/r2_flops_$c.json 2> gpurun_out/r2_flops_$c.err
  echo $c $(wc -l < gpurun_out/r2_flops_$c.csv) lines
done
