#!/bin/bash
# run a selection of GPU tests:  gpurun -- 'bash tools/gpu_one.sh <tag> <pytest args...>'
o=gpurun_out/${1:-one}; shift
mkdir -p $o
( timeout 300 python -m pytest "$@" -m gpu -q -x ) > $o/pytest.log 2>&1
tail -25 $o/pytest.log | cut -c1-400
