#!/bin/bash
o=gpurun_out/${1:-t}
mkdir -p $o
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_blocks_dense.py -m gpu -q -x -k "fused_into and (16-8-256 or 24-20-128 or 18-30)" > $o/sanitizer_memcheck_fused.log 2>&1; echo "memcheck rc=$?"
tail -6 $o/sanitizer_memcheck_fused.log | cut -c1-300
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "partial_k_chunk and (cin8 or cin96 or cin40 or h4-w4)" > $o/sanitizer_memcheck_partial.log 2>&1; echo "memcheck rc=$?"
tail -6 $o/sanitizer_memcheck_partial.log | cut -c1-300
