#!/bin/bash
# ablation of post_attn<160> (timing only, results are wrong by construction): which removed piece shortens the launch?
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ak; mkdir -p $O
L=$PWD/mapf_gpt_b200
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 2 > $O/b_$name.json 2>$O/b_$name.err
  python - <<PY
import json
f="$O/b_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), round(d['ms_per_step'],2), {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-600:])
PY
}
run base X=1
for v in GELU QKV_STORE X_STORE X_LOAD ATT_LOAD ALLMEM; do
run abl_$v MAPF_GPT_B200_LIB_PATH=$L/libabl_$v.so
done
run base2 X=1
