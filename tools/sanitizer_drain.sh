#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the bulk-store drain of the tcgen05 convolution kernels and the stem
# (round 2).  Small cases only: every launch is replayed slowly.  -> profiles/r2b_compute_sanitizer.txt
o=gpurun_out/san_r2
mkdir -p $o
run() { name=$1; tool=$2; shift 2; timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" > $o/$name.log 2>&1; echo "$name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $o/$name.log | tail -3 | cut -c1-200; }
run memcheck_drain memcheck tests/test_gpu_conv.py -m gpu -q -x -k "drain_variants and bulk"
run racecheck_drain racecheck tests/test_gpu_conv.py -m gpu -q -x -k "drain_variants and bulk and 1"
run memcheck_stem memcheck tests/test_gpu_stem.py -m gpu -q -x -k "bulk-store and (150 or 100)"
