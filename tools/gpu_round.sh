#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list of the bench, ncu --set full captures.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-a}
o=gpurun_out/$tag
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $o/pytest_gpu.log 2>&1
tail -3 $o/pytest_gpu.log
timeout 600 python bench.py --dump-profile $o/conv_profile.json > $o/bench_n1.json 2> $o/bench_n1.err
cat $o/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_ref.json 2> $o/bench_ref.err
cat $o/bench_ref.json
timeout 200 python tools/prof_vq.py > $o/vq_timing.json 2> $o/vq_timing.err
timeout 100 python tools/probe_write_bw.py > $o/write_bw.json 2>&1; cat $o/write_bw.json
cat $o/vq_timing.json
if [ "$2" != "noncu" ]; then
MCQ_CUDA_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $o/launches.csv python bench.py --steps 2 --warmup 3 > $o/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 1 -c 1 -f -o $o/conv_pair3 python tools/prof_conv.py > $o/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 4 -c 1 -f -o $o/conv_1pass python tools/prof_conv.py >> $o/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vq_assign -f -o $o/vq python tools/prof_vq.py --once > $o/ncu_vq.log 2>&1
fi
ls -la $o
