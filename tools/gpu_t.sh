#!/bin/bash
o=gpurun_out/${1:-t}
mkdir -p $o
( time timeout 500 python -m pytest tests/test_neon.py tests/test_gpu_model.py -m gpu -q -x -k "c128 or large_unaligned" ) > $o/pytest.log 2>&1
tail -30 $o/pytest.log | cut -c1-400
