#!/bin/bash
# bench-only sweep over environment knobs: tools/gpu_sweep.sh tag "VAR=a VAR=b ..."
tag=${1:-s}
o=gpurun_out/$tag
mkdir -p $o
i=0
for setting in "${@:2}"; do
  i=$((i+1))
  env $setting timeout 200 python bench.py --steps 10 --warmup 3 > $o/bench_$i.json 2> $o/bench_$i.err
  python - <<PY
import json
try:
    b=json.load(open("$o/bench_$i.json"))
    print("$setting", "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "e2e", round(b["e2e"]["value"],1), "clk", b["clocks"]["sm_mhz"])
except Exception as e:
    print("$setting", "FAILED", e)
PY
done
