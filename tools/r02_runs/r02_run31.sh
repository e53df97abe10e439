#!/bin/bash
# the SM clock post_attn and attention really run at inside a power-capped step, and the per-tile gap of a CTA slot (tools/clock_probe.py)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02aj; mkdir -p $O
for v in default "MAPF_GPT_B200_POST_CL=1" "MAPF_GPT_B200_POST_PERSIST=1"; do
  echo "== $v"; env $( [ "$v" = default ] && echo X=1 || echo $v ) timeout 300 python tools/clock_probe.py 2> $O/clock_probe.err | tee -a $O/clock_probe.txt; tail -2 $O/clock_probe.err
done
