#!/bin/bash
# quick perf iteration: conv parity tests + bench (+ per-launch profile dump)
tag=${1:-x}
o=gpurun_out/$tag
mkdir -p $o
( timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q ) > $o/pytest.log 2>&1
tail -4 $o/pytest.log
timeout 300 python bench.py --dump-profile $o/conv_profile.json > $o/bench_n1.json 2> $o/bench_n1.err
python - <<PY
import json
b=json.load(open("$o/bench_n1.json"))
print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["value"], "exec_frac", b["roofline"]["executed_frac"], "clocks", b["clocks"])
PY
tail -3 $o/bench_n1.err
