#!/bin/bash
# ncu --set full of block 0's launches (attention gathering q/k/v from the table, post_attn taking x from it) and of the 6M kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"post_attn_kernel|attn_persistent" -s 0 -c 2 -f -o gpurun_out/prof_r02_block0 \
    python bench.py --quick --envs 128 --steps 1 --warmup 1 > gpurun_out/prof_r02_block0.log 2>&1; echo "block0 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"post_attn_kernel|attn_persistent" -s 2 -c 2 -f -o gpurun_out/prof_r02_6m \
    python bench.py --quick --model 6M --map wfi_warehouse --agents 192 --envs 43 --steps 1 --warmup 1 > gpurun_out/prof_r02_6m.log 2>&1; echo "6M rc=$?"
ls -la gpurun_out/prof_r02_block0.ncu-rep gpurun_out/prof_r02_6m.ncu-rep
