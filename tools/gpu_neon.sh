#!/bin/bash
o=gpurun_out/${1:-neon}
mkdir -p $o
timeout 400 python tools/bench_neon.py --n 8 --hw 512 --dense 1 --layers 8 --steps 3 > $o/neon_a800_16_dense.txt 2> $o/neon_dense.err; tail -3 $o/neon_dense.err; cat $o/neon_a800_16_dense.txt
NEON_SIZE=16,8,4,2,2 timeout 400 python tools/bench_neon.py --n 8 --hw 512 --channel 32 --dense 1 --layers 8 --steps 3 > $o/neon_c32_dense.txt 2> $o/neon_c32.err; tail -3 $o/neon_c32.err; cat $o/neon_c32_dense.txt
