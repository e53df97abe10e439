#!/bin/bash
# A/B: MLP in 4 chunks of 256 hidden columns (WIDE, libnext.so) vs 8 chunks of 128 (libnext_narrow.so) in post_attn_kernel<256>
cd "$(dirname "$0")/../.."
O=gpurun_out/r02z; mkdir -p $O
NEW=$PWD/mapf_gpt_b200/libnext.so; OLD=$PWD/mapf_gpt_b200/libnext_narrow.so
MAPF_GPT_B200_LIB_PATH=$NEW timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests_new.log 2>&1; echo "tests(new) rc=$?"; tail -4 $O/tests_new.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b6M_$name.json 2>$O/b6M_$name.err
  python - <<PY
import json
for f in ("$O/b6M_$name.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.015}, d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run narrow MAPF_GPT_B200_LIB_PATH=$OLD
run wide MAPF_GPT_B200_LIB_PATH=$NEW
run narrowb MAPF_GPT_B200_LIB_PATH=$OLD
run wideb MAPF_GPT_B200_LIB_PATH=$NEW
run wide_oneshot MAPF_GPT_B200_LIB_PATH=$NEW MAPF_GPT_B200_POST_PERSIST=0
