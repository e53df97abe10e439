#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (table-gathering attention, LayerNorm-folded GEMM epilogues, fp32 verification
# forward, density metric, persistent post_attn<256>): smoke() + a test subset
cd "$(dirname "$0")/../.."
O=gpurun_out/r02p; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/san_smoke.txt 2>&1; echo "smoke memcheck rc=$?"; tail -3 $O/san_smoke.txt
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -x -q \
  -k "golden or fp32_verification_mode_logits or density or validated or fused_and_generic or kernel_variants_agree or max_free or reset_states" > $O/san_tests.txt 2>&1; echo "tests memcheck rc=$?"; tail -6 $O/san_tests.txt
grep -c "ERROR SUMMARY" $O/san_smoke.txt $O/san_tests.txt; grep "ERROR SUMMARY" $O/san_smoke.txt $O/san_tests.txt | tail -3
