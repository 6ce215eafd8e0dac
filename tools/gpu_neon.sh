#!/bin/bash
tag=${1:-neon}
o=gpurun_out/$tag
mkdir -p $o
timeout 300 python tools/bench_neon.py --n 2 --hw 256 --steps 2 > $o/neon_small.json 2> $o/neon_small.err; tail -3 $o/neon_small.err; cat $o/neon_small.json
timeout 400 python tools/bench_neon.py --n 8 --hw 512 > $o/neon_a800_16.json 2> $o/neon_a800_16.err; tail -5 $o/neon_a800_16.err; cat $o/neon_a800_16.json
timeout 300 python tools/bench_neon.py --n 8 --hw 512 --decode-passes 1 > $o/neon_a800_16_dec1.json 2> $o/neon_a800_16_dec1.err; tail -3 $o/neon_a800_16_dec1.err; cat $o/neon_a800_16_dec1.json
