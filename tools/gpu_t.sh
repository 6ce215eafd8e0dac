#!/bin/bash
o=gpurun_out/${1:-t}
mkdir -p $o
timeout 400 python tools/bench_compressor.py --layers 12 > $o/compressor_c192_m6.txt 2> $o/c192.err; tail -3 $o/c192.err; cat $o/compressor_c192_m6.txt
