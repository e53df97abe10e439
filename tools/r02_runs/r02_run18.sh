#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -x -q -k "validation_loss or block0 or golden" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -12 $O/tests.log
