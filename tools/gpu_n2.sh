#!/bin/bash
o=gpurun_out/${1:-n2}
mkdir -p $o
nvidia-smi -L > $o/smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err
cat $o/bench_n2.json; tail -5 $o/bench_n2.err
