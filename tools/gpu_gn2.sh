#!/bin/bash
tag=${1:-gn2}
o=gpurun_out/$tag
mkdir -p $o
for cfg in "256 4" "512 2" "512 1" "1024 1"; do
  set -- $cfg
  echo "== THREADS=$1 CTAS_PER_SM=$2"
  MCQ_GN_THREADS=$1 MCQ_GN_CTAS_PER_SM=$2 timeout 120 python -m pytest tests/test_blocks_dense.py -m gpu -q 2>&1 | tail -1
  MCQ_GN_THREADS=$1 MCQ_GN_CTAS_PER_SM=$2 timeout 120 python tools/prof_groupnorm.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        for r in json.loads(l)['groupnorm']: print(r['hw'], r['c'], r['passes'], round(r['ms']*1000,1), 'us', round(r['GBps']), 'GB/s', round(r['frac_of_measured_hbm'],3))
"
done | tee $o/gn_sweep.txt
