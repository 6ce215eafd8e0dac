#!/bin/bash
# One measurement pass over the final code of a round (outputs gpurun_out/rN_*; copy what is quoted into profiles/): full GPU suite, smoke, bench lines, ncu launch list, ncu full captures, speed protocol
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -140 > $O/rN_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/rN_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile $O/rN_conv_launch_times.json > $O/rN_bench_n1.json 2> $O/rN_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/rN_bench_reference_arm.json 2> $O/rN_ref.err
MCQ_CUDA_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/rN_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-strong --no-gpu-baseline --no-cpu-baseline --no-roofline > $O/rN_ncu_bench.log 2>&1
python tools/summarize_launches.py $O/rN_launches_bench_steps2.csv 2 $O/rN_launch_summary.json > $O/rN_launch_summary.txt 2>&1
PROF_HW=64 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_pair -s 2 -c 1 -f -o $O/rN_conv_pair3_64 python tools/prof_conv.py > $O/rN_ncu1.log 2>&1
PROF_HW=64 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_pair -s 5 -c 1 -f -o $O/rN_conv_pair1_64 python tools/prof_conv.py > $O/rN_ncu2.log 2>&1
timeout 300 python tools/prof_latency.py > $O/rN_latency_small_layers.txt 2>&1
timeout 300 python tools/prof_stem.py > $O/rN_stem_timing.txt 2>&1
timeout 600 python -m mcquic_b200 --speed -qp 2 --synthetic > $O/rN_speed_qp2.txt 2>&1
timeout 600 python -m mcquic_b200 --speed -qp 1 --synthetic > $O/rN_speed_qp1.txt 2>&1
tail -6 $O/rN_pytest_gpu.txt | cut -c1-200
cat $O/rN_smoke.txt | tail -3
tail -3 $O/rN_bench_n1.err
cut -c1-900 $O/rN_bench_n1.json; echo
cut -c1-300 $O/rN_bench_reference_arm.json; echo
tail -14 $O/rN_launch_summary.txt
tail -12 $O/rN_latency_small_layers.txt | head -11
tail -2 $O/rN_stem_timing.txt; tail -n 2 $O/rN_speed_qp2.txt; tail -n 2 $O/rN_speed_qp1.txt
