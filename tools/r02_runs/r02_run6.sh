#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r02f; mkdir -p $O
run() { name=$1; shift
  env "$@" MAPF_GPT_B200_DEBUG_OCC=1 timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$name.json 2>$O/b2M_$name.err
  grep post_attn $O/b2M_$name.err
  python - <<PY
import json
for f in ("$O/b2M_$name.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['kernels']['attention']['avg_ms'], d['kernels']['post_attn_fused']['avg_ms'], d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run g296 X=1
run g148 MAPF_GPT_B200_POST_GRID=148
run g222 MAPF_GPT_B200_POST_GRID=222
run g592 MAPF_GPT_B200_POST_GRID=592
run cl1_persist MAPF_GPT_B200_POST_CL=1
run cl1_oneshot MAPF_GPT_B200_POST_CL=1 MAPF_GPT_B200_POST_PERSIST=0
