#!/bin/bash
# quick check that bench.py still emits exactly one valid JSON line (after touching its reporting code)
o=gpurun_out/${1:-bc}
mkdir -p $o
timeout 300 python bench.py --steps 3 --warmup 3 > $o/bench_n1.json 2> $o/bench_n1.err
echo "stdout lines: $(wc -l < $o/bench_n1.json)"; tail -2 $o/bench_n1.err
python -c "import json; d=json.load(open('$o/bench_n1.json')); print('ok', d['value'], d['roofline']['conv_share_of_step']); [print(k, v) for k, v in d['roofline']['by_layer_class'].items()]"
