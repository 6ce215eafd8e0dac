#!/bin/bash
# profile refresh: launch list of the bench + ncu --set full of the dominant kernels
tag=${1:-p}
o=gpurun_out/$tag
mkdir -p $o
MCQ_CUDA_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $o/launches.csv python bench.py --steps 2 --warmup 3 > $o/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 1 -c 1 -f -o $o/conv_pair3 python tools/prof_conv.py > $o/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 4 -c 1 -f -o $o/conv_pair1 python tools/prof_conv.py >> $o/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem" -f -o $o/misc python tools/prof_misc.py > $o/ncu_misc.log 2>&1
tail -2 $o/ncu_misc.log
ls -la $o
