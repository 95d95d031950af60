#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_rate_def_types.py tests/test_rate_model.py tests/test_def_types.py -m gpu -q ) > gpurun_out/r2ac_new.log 2>&1; tail -n 40 gpurun_out/r2ac_new.log | cut -c1-300
