#!/bin/bash
# persistent post_attn<256> (6M): also prefetch the next att tile into L2 (-DMG_POST_PF_ATT)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02at; mkdir -p $O
L=$PWD/mapf_gpt_b200
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b_$name.json 2>$O/b_$name.err
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
run base_$rep X=1
run pfatt_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_pfatt.so
run nopersist_$rep MAPF_GPT_B200_POST_PERSIST=0
done
