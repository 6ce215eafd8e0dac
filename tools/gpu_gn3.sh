#!/bin/bash
o=gpurun_out/${1:-gnf}
mkdir -p $o
timeout 200 python tools/prof_gn_fused.py > $o/gn_fused_timing.json 2> $o/gn_fused.err; tail -3 $o/gn_fused.err; cat $o/gn_fused_timing.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 2 -c 1 -f -o $o/gn_apply python tools/prof_gn_fused.py > $o/ncu_gn_apply.log 2>&1; tail -2 $o/ncu_gn_apply.log
