#!/bin/bash
# compute-sanitizer memcheck over the last session's kernels / paths: stream lanes, tail-aware stores, 24-bit residual stream
# (post_attn and the pair GEMM epilogue), embed_tile_kernel, successor L2 prefetch
cd "$(dirname "$0")/../.."
O=gpurun_out/r02az; mkdir -p $O
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -x -q \
  -k "lanes_are_bit or tail_aware or 24bit or embed_tile" > $O/san_tests3.txt 2>&1; echo "tests memcheck rc=$?"; tail -5 $O/san_tests3.txt
grep "ERROR SUMMARY" $O/san_tests3.txt | tail -3
