#!/bin/bash
o=gpurun_out/${1:-t}
mkdir -p $o
( time timeout 900 python -m pytest tests/test_neon.py -m gpu -q -x ) > $o/pytest.log 2>&1
tail -5 $o/pytest.log | cut -c1-600
NEON_SIZE=16,8,4,2,2 timeout 400 python tools/bench_neon.py --n 8 --hw 512 --channel 32 --dense 1 --layers 8 --steps 3 > $o/neon_c32_dense.txt 2> $o/neon_c32.err; tail -3 $o/neon_c32.err; cat $o/neon_c32_dense.txt
timeout 400 python tools/bench_neon.py --n 8 --hw 512 --layers 24 --steps 5 > $o/neon_a800_16.txt 2> $o/neon_a800.err; tail -3 $o/neon_a800.err; head -8 $o/neon_a800_16.txt
timeout 400 python tools/bench_neon.py --n 8 --hw 512 --decode-passes 1 --steps 5 > $o/neon_a800_16_dec1.txt 2> $o/neon_a800.err; cat $o/neon_a800_16_dec1.txt
