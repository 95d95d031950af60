#!/bin/bash
# Round-2 session AA: Barlat in the def-type / rate kernels, Hosford-exponent column from the def-type kernel; whole suite.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_def_types.py tests/test_rate_model.py -m gpu -q ) > gpurun_out/r2aa_new.log 2>&1; tail -n 30 gpurun_out/r2aa_new.log
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2aa_pytest.log 2>&1; tail -n 5 gpurun_out/r2aa_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
