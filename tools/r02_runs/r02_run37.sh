#!/bin/bash
# persistent post_attn<160> again on the final build (24-bit residual, smaller loads): does the tile loop win now?
cd "$(dirname "$0")/../.."
O=gpurun_out/r02aq; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 $ARGS > $O/b_$name.json 2>$O/b_$name.err
  python - <<PY
import json
f="$O/b_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), d['roofline']['whole_step_frac'], round(d['ms_per_step'],2), {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-600:])
PY
}
for rep in 1 2; do
run default_$rep X=1
run persist_$rep MAPF_GPT_B200_POST_PERSIST=1
run lanes296_$rep MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296
done
