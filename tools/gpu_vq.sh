#!/bin/bash
tag=${1:-b}
o=gpurun_out/$tag
mkdir -p $o
( timeout 300 python -m pytest tests/test_gpu_vq.py -x -q ) > $o/pytest_vq.log 2>&1
tail -15 $o/pytest_vq.log
timeout 120 python tools/prof_vq.py > $o/vq_timing.json 2> $o/vq_timing.err
cat $o/vq_timing.json; tail -3 $o/vq_timing.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_fused -f -o $o/vq_fused python tools/prof_vq.py --once > $o/ncu_vq.log 2>&1
tail -2 $o/ncu_vq.log
