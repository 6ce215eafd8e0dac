#!/bin/bash
o=gpurun_out/${1:-neon}
mkdir -p $o
timeout 400 python tools/bench_neon.py --n 8 --hw 512 --layers 24 --steps 3 > $o/neon_layers.txt 2> $o/neon_layers.err; tail -3 $o/neon_layers.err; cat $o/neon_layers.txt
