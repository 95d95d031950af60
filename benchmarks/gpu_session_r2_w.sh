#!/bin/bash
# Round-2 session W: Yld2004-18p (barlat) K1 / history / K2 parity, then the whole GPU suite + smoke.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_barlat.py -m gpu -q -x ) > gpurun_out/r2w_barlat.log 2>&1; tail -n 30 gpurun_out/r2w_barlat.log
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2w_pytest.log 2>&1; tail -n 6 gpurun_out/r2w_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
