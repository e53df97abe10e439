#!/bin/bash
# GELU micro-variants of post_attn (clamp on x^2; 1/8 and 2/8 of the pairs on the FMA-only form) and stream lanes with the full
# attention grid (tail overlap only), 2M and 6M; every leg twice, interleaved
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ae; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 $ARGS > $O/b_$name.json 2>$O/b_$name.err
  python - <<PY
import json
f="$O/b_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), d['roofline']['whole_step_frac'], round(d['ms_per_step'],2), {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
L=$PWD/mapf_gpt_b200
for rep in 1 2; do
ARGS=""
run base_$rep X=1
run clampt_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_clampt.so
run mix1_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_mix1.so
run mix2_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_mix2.so
run lanes296_$rep MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296
run lanes296_c4096_$rep MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296 MAPF_GPT_B200_CHUNK_SEQS=4096
ARGS="--model 6M --map wfi_warehouse --agents 192 --envs 512 --steps 4"
run 6M_base_$rep X=1
run 6M_lanes296_$rep MAPF_GPT_B200_LANES=2 MAPF_GPT_B200_LANE_ATTN_GRID=296
done
