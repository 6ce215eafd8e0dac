#!/bin/bash
# compute-sanitizer over the simple kernels and the new tcgen05 paths (small cases only: every launch is replayed slowly).
# Usage: gpurun --timeout 600 -- 'bash tools/gpu_sanitizer.sh <tag>'   -> profiles/r1e_compute_sanitizer.txt
o=gpurun_out/${1:-san}
mkdir -p $o
run() { name=$1; tool=$2; shift 2; timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" > $o/$name.log 2>&1; echo "$name rc=$?"; tail -4 $o/$name.log | cut -c1-200; }
run memcheck_simple memcheck tests/test_blocks_dense.py tests/test_neon.py -m gpu -q -x -k "groupnorm_kernel or add_scaled or groupnorm_errors"
run racecheck_groupnorm racecheck tests/test_blocks_dense.py -m gpu -q -x -k "groupnorm_kernel and (64-64 or 7-9 or 4-4)"
run memcheck_fused memcheck tests/test_blocks_dense.py -m gpu -q -x -k "fused_into and (16-8-256 or 24-20-128 or 18-30)"
run memcheck_partial memcheck tests/test_gpu_conv.py -m gpu -q -x -k "partial_k_chunk and (cin8 or cin96 or cin40 or h4-w4)"
