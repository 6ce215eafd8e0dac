#!/bin/bash
# end-of-session evidence: smoke(), launch list of the bench, ncu --set full of the dominant kernels, VQ timing
tag=${1:-z}
o=gpurun_out/$tag
mkdir -p $o
timeout 300 python __graft_entry__.py --smoke > $o/smoke.log 2>&1; tail -2 $o/smoke.log
timeout 300 python bench.py --dump-profile $o/conv_profile.json > $o/bench_n1.json 2> $o/bench_n1.err; cut -c1-400 $o/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_ref.json 2> $o/bench_ref.err
MCQ_CUDA_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $o/launches.csv python bench.py --steps 2 --warmup 3 > $o/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 1 -c 1 -f -o $o/conv_pair3 python tools/prof_conv.py > $o/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 4 -c 1 -f -o $o/conv_pair1 python tools/prof_conv.py >> $o/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_fused -f -o $o/vq_fused python tools/prof_vq.py --once > $o/ncu_vq.log 2>&1
timeout 200 python tools/prof_vq.py > $o/vq_timing.json 2> $o/vq_timing.err; cat $o/vq_timing.json
timeout 200 python tools/prof_latency.py > $o/latency.txt 2>&1; tail -3 $o/latency.txt
ls $o
