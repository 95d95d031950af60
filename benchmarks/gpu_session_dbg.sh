#!/bin/bash
python benchmarks/debug_rate_dt.py 2>&1 | tail -6 | cut -c1-300
( timeout 900 python -m pytest tests/test_rate_def_types.py -m gpu -q ) 2>&1 | tail -4
