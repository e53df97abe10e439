#!/bin/bash
# stream lanes (attention of one chunk overlapping post_attn of the other, 2M): bit-identity test + A/B over the attention grid
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ad; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "lanes" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$name.json 2>$O/b2M_$name.err
  python - <<PY
import json
f="$O/b2M_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), d['roofline']['whole_step_frac'], round(d['ms_per_step'],2), {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d.get('stream_lanes',{}).get('single_lane_ms_per_step_with_kernel_events'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run lanes1 MAPF_GPT_B200_LANES=1
run lanes2_g148 MAPF_GPT_B200_LANES=2
run lanes2_g296 MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296
run lanes2_g222 MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=222
run lanes2_g74 MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=74
run lanes2_g148_c4096 MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_CHUNK_SEQS=4096
run lanes1b MAPF_GPT_B200_LANES=1
run lanes2_g148b MAPF_GPT_B200_LANES=2
