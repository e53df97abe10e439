#!/bin/bash
# post_attn: residual added in the LN2 epilogue instead of pre-loaded into the accumulator (-DMG_POST_XLATE=1): tests + A/B, both tile-loop modes
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ar; mkdir -p $O
L=$PWD/mapf_gpt_b200
MAPF_GPT_B200_LIB_PATH=$L/libvar_xlate.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -x -q -k "not stress and not two_devices" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
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
ARGS=""
run base_$rep X=1
run xlate_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_xlate.so
run xlate_persist_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_xlate.so MAPF_GPT_B200_POST_PERSIST=1
ARGS="--model 6M --map wfi_warehouse --agents 192 --envs 512 --steps 4"
run 6M_base_$rep X=1
run 6M_xlate_$rep MAPF_GPT_B200_LIB_PATH=$L/libvar_xlate.so
done
