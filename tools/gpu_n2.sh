#!/bin/bash
o=gpurun_out/${1:-n2}
mkdir -p $o
timeout 300 python bench.py --steps 5 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err; cut -c1-300 $o/bench_n1.json; tail -3 $o/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err
cut -c1-700 $o/bench_n2.json; tail -3 $o/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $o/bench_ref_n2.json 2> $o/bench_ref_n2.err
cut -c1-200 $o/bench_ref_n2.json
