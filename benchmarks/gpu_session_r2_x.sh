#!/bin/bash
# Round-2 session X: Yld2004-18p element blocks, cmadx_sym3_eigh, cmadx_mp_model_partials: parity; then the whole suite.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_barlat.py tests/test_sym3_eigh.py tests/test_model_partials.py -m gpu -q ) > gpurun_out/r2x_new.log 2>&1; tail -n 40 gpurun_out/r2x_new.log
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2x_pytest.log 2>&1; tail -n 6 gpurun_out/r2x_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
