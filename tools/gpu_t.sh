#!/bin/bash
o=gpurun_out/${1:-t}
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/pytest_gpu.log 2>&1
tail -6 $o/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --dump-profile $o/conv_profile.json > $o/bench_n1.json 2> $o/bench_n1.err; tail -3 $o/bench_n1.err
cat $o/bench_n1.json
