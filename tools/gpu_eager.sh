#!/bin/bash
o=gpurun_out/${1:-eager}
mkdir -p $o
timeout 300 python tests/measure_eager_gpu.py 64 > $o/eager_q1_n64.json 2> $o/eager_q1.err; tail -2 $o/eager_q1.err; cat $o/eager_q1_n64.json
timeout 400 python tests/measure_eager_gpu.py neon 8 512 > $o/eager_neon.json 2> $o/eager_neon.err; tail -2 $o/eager_neon.err; cat $o/eager_neon.json
