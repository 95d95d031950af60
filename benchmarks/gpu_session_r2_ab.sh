#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_def_types.py tests/test_barlat.py -m gpu -q ) > gpurun_out/r2ab_new.log 2>&1; tail -n 12 gpurun_out/r2ab_new.log
