#!/bin/bash
o=gpurun_out/${1:-exp}
mkdir -p $o
for p in 2 4; do timeout 200 python tools/exp_halves.py $p 2>&1 | tail -2; done | tee $o/exp_halves.txt
