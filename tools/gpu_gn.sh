#!/bin/bash
# GPU visit for the dense-norm blocks: parity tests of the new path, mcq_groupnorm timing, one ncu --set full capture.
tag=${1:-gn}
o=gpurun_out/$tag
mkdir -p $o
( time timeout 300 python -m pytest tests/test_blocks_dense.py -m gpu -q ) > $o/pytest_gn.log 2>&1
tail -25 $o/pytest_gn.log
timeout 120 python tools/prof_groupnorm.py > $o/gn_timing.json 2> $o/gn_timing.err
cat $o/gn_timing.json; tail -3 $o/gn_timing.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:groupnorm -s 2 -c 1 -f -o $o/gn python tools/prof_groupnorm.py > $o/ncu_gn.log 2>&1
tail -2 $o/ncu_gn.log
ls -la $o
