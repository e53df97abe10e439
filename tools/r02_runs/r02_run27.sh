#!/bin/bash
# sequences per chunk as a whole number of post_attn waves (148 sequences = one wave of 296 resident CTAs): 8288 = 56 waves, ...
cd "$(dirname "$0")/../.."
O=gpurun_out/r02af; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 $ARGS > $O/b_$name.json 2>$O/b_$name.err
  python - <<PY
import json
f="$O/b_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), d['roofline']['whole_step_frac'], round(d['ms_per_step'],2), {k:(v['avg_ms'],v['launches']) for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
for rep in 1 2; do
ARGS=""
run c8192_$rep X=1
run c8288_$rep MAPF_GPT_B200_CHUNK_SEQS=8288
run c9472_$rep MAPF_GPT_B200_CHUNK_SEQS=9472
run c13172_$rep MAPF_GPT_B200_CHUNK_SEQS=13172
run c16428_$rep MAPF_GPT_B200_CHUNK_SEQS=16428
run c8288_lanes_$rep MAPF_GPT_B200_CHUNK_SEQS=8288 MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296
done
