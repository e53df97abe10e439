#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU): launch list of a short bench run + one --set full capture per hot kernel,
# for the 2M (fused) path and the 85M (generic) path.  Numbers printed by bench.py under ncu are never bench values.
set -u
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --quick --steps 2 --warmup 1 > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"post_attn_kernel|attn_persistent|block0_lookup|last_attn" -s 2 -c 6 -f -o gpurun_out/prof_$tag \
    python bench.py --quick --envs 128 --steps 1 --warmup 1 > gpurun_out/prof_$tag.log 2>&1
echo "full capture (2M) rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${tag}_85m.csv \
    python bench.py --quick --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 --steps 1 --warmup 1 > gpurun_out/launches_${tag}_85m.log 2>&1
echo "launch list (85M) rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_pair_persistent|attn_ts_kernel|embed_kernel" -s 7 -c 7 -f -o gpurun_out/prof_${tag}_85m \
    python bench.py --quick --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 --steps 1 --warmup 1 > gpurun_out/prof_${tag}_85m.log 2>&1
echo "full capture (85M) rc=$?"
ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/prof_${tag}_85m.ncu-rep
