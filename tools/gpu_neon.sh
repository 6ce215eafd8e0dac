#!/bin/bash
tag=${1:-neon}
o=gpurun_out/$tag
mkdir -p $o
( time timeout 400 python -m pytest tests/test_neon.py tests/test_blocks_dense.py -m gpu -q ) > $o/pytest_neon.log 2>&1
tail -40 $o/pytest_neon.log
